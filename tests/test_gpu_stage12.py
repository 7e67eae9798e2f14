"""GPU parity of stage 1 / 1b / 1' / 2 (similarity maps, fused up-sample + softmax + max + fusion, threshold glue)
through the C ABI, against (a) the golden bundles recorded from the live reference and (b) the numpy oracle on larger
seeded inputs.  Float outputs: abs 2e-6; labels: identical wherever the top-2 margin >= 1e-5;
everything downstream of a given label map (fusion, glue) is integer work and must be exactly equal."""
import numpy as np
import pytest
import torch

from oracle import css_oracle as O
from tests.helpers import load_golden, top2_margin, assert_labels_match

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("name", ["stage12_mix_c21", "stage12_cross_c19", "stage12_ori_c21"])
def test_stage12_against_reference_golden(name):
    import css_b200
    g = load_golden(name)
    kind = str(g["kind"])
    H, W, C, temp = int(g["H"]), int(g["W"]), int(g["C"]), float(g["temp"])
    protos = dev(g["prototypes"])
    sim = css_b200.ops.cos_sim_map(dev(g["rep_u"]), protos)
    np.testing.assert_allclose(sim.cpu().numpy(), g["sim"], rtol=0, atol=1e-6)
    m_cls = top2_margin(O.softmax(g["pred_u_large_raw"])) if kind == "ori" else \
        top2_margin(O.softmax(O.bilinear_upsample(g["pred_u"], (H, W))))
    if kind == "ori":
        conf, lab = css_b200.ops.cls_pseudo_label(dev(g["pred_u"]), (H, W))
        assert lab.dtype == torch.int64 and conf.dtype == torch.float32
        np.testing.assert_allclose(conf.cpu().numpy(), g["conf_cls"], rtol=0, atol=2e-6)
        assert_labels_match(lab.cpu().numpy(), g["label_cls"], m_cls, "label_cls")
        return
    m_rep = top2_margin(O.softmax(g["sim_large"] / np.float32(temp)))
    o = css_b200.ops.pseudo_labels(dev(g["rep_u"]), dev(g["pred_u"]), protos, temp, (H, W),
                                   fuse="mix" if kind == "mix" else "none")
    np.testing.assert_allclose(o["conf_cls"].cpu().numpy(), g["conf_cls"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(o["conf_rep"].cpu().numpy(), g["conf_rep"], rtol=0, atol=2e-6)
    prob_all = css_b200.ops.proto_softmax_sim(dev(g["rep_all"]), protos, temp)
    np.testing.assert_allclose(prob_all.cpu().numpy(), g["prob_all"], rtol=0, atol=1e-6)
    if kind == "cross":
        assert_labels_match(o["label_cls"].cpu().numpy(), g["label_cls"], m_cls, "label_cls")
        assert_labels_match(o["label_rep"].cpu().numpy(), g["label_rep"], m_rep, "label_rep")
    else:
        fused = o["fused"].cpu().numpy()
        assert fused.dtype == np.float32
        bad = fused != g["fused"]
        assert (np.minimum(m_cls, m_rep)[bad] < 1e-5).all()
        # fusion given the kernel's own labels is exact integer work
        assert np.array_equal(fused, O.mix_fuse(o["label_cls"].cpu().numpy(), o["label_rep"].cpu().numpy(), C))


@pytest.mark.parametrize("B,C,h,w,H,W,temp", [(2, 21, 81, 81, 321, 321, 0.5), (1, 19, 49, 97, 193, 385, 0.5),
                                               (2, 21, 17, 23, 50, 70, 0.1), (1, 3, 5, 4, 5, 4, 0.25),
                                               (1, 32, 6, 6, 1, 1, 0.5),
                                               (1, 21, 30, 30, 30, 30, 0.5),      # staged tile > 48 KB (opt-in smem)
                                               (1, 21, 64, 64, 16, 16, 0.5),      # strong down-sampling: un-staged taps
                                               (1, 20, 12, 9, 47, 35, 0.3)])      # generic class count, IEEE division by temp
def test_stage12_against_oracle(B, C, h, w, H, W, temp):
    import css_b200
    from css_b200 import synth
    t = synth.teacher_batch(B, C, h, w, seed=5 + C + h)
    protos = 0.5 * t["centers"] + 0.3 * synth.warm_prototypes(C, seed=9)
    protos[C // 2] = 0                                           # a never-touched prototype row stays zero -> sim 0
    rep_u, pred_u = t["rep_u"].numpy(), t["pred_u"].numpy()
    o = css_b200.ops.pseudo_labels(t["rep_u"].cuda(), t["pred_u"].cuda(), protos.cuda(), temp, (H, W), fuse="mix")
    sim_ref = O.cos_sim_map(rep_u, protos.numpy())
    sim = css_b200.ops.cos_sim_map(t["rep_u"].cuda(), protos.cuda()).cpu().numpy()
    np.testing.assert_allclose(sim, sim_ref, rtol=0, atol=1e-6)
    assert np.all(sim[:, C // 2] == 0)
    # bilinear + softmax + max fed the kernel's own low-res map: the up-sampling arithmetic is op-for-op the oracle's
    up = O.bilinear_upsample(sim, (H, W))
    conf_rep, label_rep = O.softmax_max(up, temp)
    conf_cls, label_cls = O.cls_pseudo_label(pred_u, (H, W))
    np.testing.assert_allclose(o["conf_rep"].cpu().numpy(), conf_rep, rtol=0, atol=2e-6)
    np.testing.assert_allclose(o["conf_cls"].cpu().numpy(), conf_cls, rtol=0, atol=2e-6)
    n1 = assert_labels_match(o["label_rep"].cpu().numpy(), label_rep, top2_margin(O.softmax(up / np.float32(temp))), "label_rep")
    n2 = assert_labels_match(o["label_cls"].cpu().numpy(), label_cls,
                             top2_margin(O.softmax(O.bilinear_upsample(pred_u, (H, W)))), "label_cls")
    assert n1 + n2 <= max(2, int(1e-4 * B * H * W))
    assert np.array_equal(o["fused"].cpu().numpy(),
                          O.mix_fuse(o["label_cls"].cpu().numpy(), o["label_rep"].cpu().numpy(), C))
    prob = css_b200.ops.proto_softmax_sim(t["rep_u"].cuda(), protos.cuda(), temp).cpu().numpy()
    np.testing.assert_allclose(prob, O.proto_softmax_sim(rep_u, protos.numpy(), temp), rtol=0, atol=1e-6)
    np.testing.assert_allclose(prob.sum(1), 1.0, atol=1e-5)


def test_zero_norm_pixels_and_single_ops():
    import css_b200
    rep = torch.randn(1, 256, 4, 5)
    rep[0, :, 1, 2] = 0                                          # zero vector: F.normalize eps keeps it at 0 -> sim 0
    protos = torch.randn(7, 256)
    sim = css_b200.ops.cos_sim_map(rep.cuda(), protos.cuda()).cpu().numpy()
    assert np.all(sim[0, :, 1, 2] == 0) and np.isfinite(sim).all()
    np.testing.assert_allclose(sim, O.cos_sim_map(rep.numpy(), protos.numpy()), atol=1e-6)
    conf, lab = css_b200.ops.rep_pseudo_label(rep.cuda(), protos.cuda(), 0.5, (13, 17))
    c2, l2, _ = O.rep_pseudo_label(rep.numpy(), protos.numpy(), 0.5, (13, 17))
    np.testing.assert_allclose(conf.cpu().numpy(), c2, atol=2e-6)
    a = torch.randint(0, 7, (2, 9, 9)).cuda()
    b = torch.randint(0, 7, (2, 9, 9)).cuda()
    assert np.array_equal(css_b200.ops.mix_fuse(a, b, 7).cpu().numpy(), O.mix_fuse(a.cpu().numpy(), b.cpu().numpy(), 7))


@pytest.mark.parametrize("name", ["glue_mix", "glue_cross", "glue_ori"])
def test_threshold_glue_exact_against_reference_golden(name):
    import css_b200
    g = load_golden(name)
    h, w = int(g["h"]), int(g["w"])
    label_all, mask_all = css_b200.ops.threshold_glue(dev(g["train_l_label"].astype(np.int64)),
                                                      dev(g["u_label"].astype(np.int64)), dev(g["conf"]), float(g["weak"]),
                                                      int(g["C"]), (h, w), str(g["strategy"]))
    assert np.array_equal(label_all.cpu().numpy(), g["label_all"].astype(np.float32))
    assert np.array_equal(mask_all.cpu().numpy(), g["mask_all"].astype(np.float32))


@pytest.mark.parametrize("strategy,B,C,H,W,h,w", [("mix", 2, 21, 321, 321, 81, 81), ("cross", 1, 19, 100, 75, 33, 20),
                                                   ("ori", 2, 21, 64, 64, 64, 64)])
def test_threshold_glue_exact_against_oracle(strategy, B, C, H, W, h, w):
    import css_b200
    from css_b200 import synth
    g = synth._gen(3)
    ll = synth.class_map(B, C, H, W, g, ignore_frac=0.1)
    lu = synth.class_map(B, C, H, W, g, ignore_frac=0.2)
    conf = torch.floor(torch.rand(B, H, W, generator=g) * 255) / 255
    la, ma = css_b200.ops.threshold_glue(ll.cuda(), lu.cuda(), conf.cuda(), 0.7, C, (h, w), strategy)
    la_ref, ma_ref = O.threshold_glue(ll.numpy(), lu.numpy(), conf.numpy(), 0.7, C, (h, w), strategy)
    assert np.array_equal(la.cpu().numpy(), la_ref) and np.array_equal(ma.cpu().numpy(), ma_ref)


@pytest.mark.parametrize("B,C,h,w", [(8, 21, 81, 81), (4, 19, 193, 193), (3, 32, 20, 21), (1, 5, 7, 9), (2, 21, 33, 31)])
def test_tensor_core_rep_pass_matches_oracle_and_fma_path(B, C, h, w):
    """css_rep_pass has two multiply sides for fp32 maps: packed FFMA2 (default) and tcgen05.mma kind::tf32 with TMEM
    accumulators and an fp32-exact hi/lo operand split (css_sim_tc.cu, opt-in).  Both must meet the same tolerances; the
    pixel-major rows are a bit-exact copy of the map on either path."""
    import css_b200
    from css_b200 import _lib, synth
    lib = _lib.load()
    d = synth.student_batch(B, C, h, w, seed=7 + C, block=4)
    protos = 0.5 * d["centers"] + 0.3 * synth.warm_prototypes(C, seed=3)
    protos[C // 2] = 0
    rep = d["rep"].cuda()
    sim_ref = O.cos_sim_map(d["rep"].numpy(), protos.numpy())
    prob_ref = O.proto_softmax_sim(d["rep"].numpy(), protos.numpy(), 0.5)
    got = {}
    try:
        for name, flag in (("fma", 0), ("tc", 1)):
            lib.css_set_rep_pass_path(flag)
            sim = css_b200.ops.cos_sim_map(rep, protos.cuda())
            prob = css_b200.ops.proto_softmax_sim(rep, protos.cuda(), 0.5)
            got[name] = (sim.cpu().numpy(), prob.cpu().numpy(), prob._css_rows.rows, prob._css_rows.norms)
    finally:
        lib.css_set_rep_pass_path(-1)
    for name, (sim, prob, rows, norms) in got.items():
        np.testing.assert_allclose(sim, sim_ref, rtol=0, atol=1e-6, err_msg=name)
        np.testing.assert_allclose(prob, prob_ref, rtol=0, atol=1e-6, err_msg=name)
        assert np.all(sim[:, C // 2] == 0), name
        assert torch.equal(rows, rep.permute(0, 2, 3, 1).reshape(-1, 256)), name
        np.testing.assert_allclose(norms.cpu().numpy(), np.linalg.norm(d["rep"].numpy(), axis=1).reshape(-1), rtol=1e-6, err_msg=name)
