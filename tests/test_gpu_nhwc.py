"""Channels-last representation maps (extension): a map produced by a `torch.channels_last` network is, in
memory, the pixel-major row table itself.  The channels-last kernels (css_rep_pass_nhwc: 2-D TMA tiles + tcgen05.mma kind::tf32
with an exact hi/lo split + TMEM accumulators; css_grad_scatter_nhwc) must give the results of the NCHW path on the same values:
similarities / probabilities abs 1e-6 against the oracle, selection exact, loss / gradient / prototypes rel 1e-4."""
import numpy as np
import pytest
import torch

from oracle import css_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def cl(t):
    return t.cuda().contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("B,C,h,w", [(8, 21, 81, 81), (2, 19, 193, 193), (3, 32, 20, 21), (1, 5, 7, 9), (2, 21, 33, 31)])
def test_similarity_maps_from_channels_last(B, C, h, w):
    import css_b200
    from css_b200 import synth
    d = synth.student_batch(B, C, h, w, seed=11 + C, block=4)
    protos = 0.5 * d["centers"] + 0.3 * synth.warm_prototypes(C, seed=3)
    protos[C // 2] = 0
    rep = cl(d["rep"])
    assert css_b200.ops.is_channels_last(rep)
    sim = css_b200.ops.cos_sim_map(rep, protos.cuda())
    prob = css_b200.ops.proto_softmax_sim(rep, protos.cuda(), 0.5)
    assert sim.is_contiguous() and prob.is_contiguous() and sim.shape == (B, C, h, w)
    np.testing.assert_allclose(sim.cpu().numpy(), O.cos_sim_map(d["rep"].numpy(), protos.numpy()), rtol=0, atol=1e-6)
    np.testing.assert_allclose(prob.cpu().numpy(), O.proto_softmax_sim(d["rep"].numpy(), protos.numpy(), 0.5), rtol=0, atol=1e-6)
    assert np.all(sim.cpu().numpy()[:, C // 2] == 0)
    cache = prob._css_rows
    assert cache.rows.data_ptr() == rep.data_ptr()                  # the map is its own row table: nothing was copied
    assert torch.equal(cache.rows, d["rep"].cuda().permute(0, 2, 3, 1).reshape(-1, 256))
    np.testing.assert_allclose(cache.norms.cpu().numpy(), np.linalg.norm(d["rep"].numpy(), axis=1).reshape(-1), rtol=1e-6)
    np.testing.assert_allclose(css_b200.ops.rep_norms_nhwc(rep).cpu().numpy(), cache.norms.cpu().numpy(), rtol=0, atol=0)
    # teacher block (similarity + fused up-sample / labels) on a channels-last map == the NCHW path on the same values
    if h * 4 <= 400:
        H, W = 4 * h - 3, 4 * w - 3
        logits = torch.randn(B, C, h, w, generator=torch.Generator().manual_seed(1)).cuda()
        a = css_b200.ops.pseudo_labels(rep, logits, protos.cuda(), 0.5, (H, W), fuse="mix")
        b = css_b200.ops.pseudo_labels(d["rep"].cuda(), logits, protos.cuda(), 0.5, (H, W), fuse="mix")
        np.testing.assert_allclose(a["conf_rep"].cpu().numpy(), b["conf_rep"].cpu().numpy(), rtol=0, atol=2e-6)
        assert (a["label_rep"] != b["label_rep"]).float().mean().item() < 1e-3
        assert torch.equal(a["label_cls"], b["label_cls"])


@pytest.mark.parametrize("B2,C,h,w,Q,Nn", [(2, 21, 81, 81, 256, 512), (3, 7, 20, 21, 16, 40)])
def test_contrast_loss_on_channels_last_map(B2, C, h, w, Q, Nn):
    import css_b200
    from css_b200 import synth
    d = synth.student_batch(B2, C, h, w, seed=5, strategy="mix", block=4)
    protos0 = synth.warm_prototypes(C, seed=6, zero_rows=(1,))
    kw = dict(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.8, alpha=0.99)
    label, mask = d["label"].cuda(), d["mask"].cuda()
    res = {}
    for name, rep0 in (("nchw", d["rep"].cuda()), ("nhwc", cl(d["rep"]))):
        crit = css_b200.Contrast_Loss(seed=9, **kw).cuda()
        protos = protos0.clone().cuda()
        prob = css_b200.ops.proto_softmax_sim(rep0, protos, 0.5)
        rep = rep0.clone().requires_grad_(True)                      # what DDP hands the loss: equal content, another address
        assert rep.stride() == rep0.stride()
        loss = crit(rep, label, mask, prob, protos)
        loss.backward()
        assert crit.last["rows_cache_mode"] == "verify" and int(crit.last["ws"].meta[css_b200._lib.META_ROWS_STALE].item()) == 0
        res[name] = dict(loss=loss.item(), grad=rep.grad, protos=protos, sel=crit.selection(), crit=crit, prob=prob)
    g_nhwc = res["nhwc"]["grad"]
    assert g_nhwc.stride() == cl(d["rep"]).stride()                  # the gradient keeps the map's memory format
    np.testing.assert_allclose(res["nhwc"]["loss"], res["nchw"]["loss"], rtol=1e-5)
    torch.testing.assert_close(res["nhwc"]["protos"], res["nchw"]["protos"], rtol=1e-5, atol=1e-6)
    assert res["nhwc"]["sel"]["present"] == res["nchw"]["sel"]["present"]
    for k in range(res["nhwc"]["sel"]["V"]):
        assert np.array_equal(res["nhwc"]["sel"]["valid_ids"][k], res["nchw"]["sel"]["valid_ids"][k])
        assert np.array_equal(res["nhwc"]["sel"]["hard_ids"][k], res["nchw"]["sel"]["hard_ids"][k])
    gn, gc = g_nhwc.contiguous().cpu().numpy(), res["nchw"]["grad"].cpu().numpy()
    assert np.array_equal(gn != 0, gc != 0)
    np.testing.assert_allclose(gn, gc, rtol=1e-4, atol=2e-5 * np.abs(gc).max())
    # and against the oracle (device draws fed back)
    crit = res["nhwc"]["crit"]
    sel = res["nhwc"]["sel"]
    a, n = crit.sample_indices(9, 0)
    slots = [k for k in range(sel["V"]) if sel["n_hard"][k] > 0]
    sampler = O.RecordedDraws([a.cpu().numpy()[k] for k in slots], [n.cpu().numpy()[k].reshape(-1) for k in slots])
    p_or = protos0.numpy().copy()
    l_or, g_or, info = O.contrast_loss(d["rep"].numpy(), d["label"].numpy(), d["mask"].numpy(), res["nhwc"]["prob"].cpu().numpy(), p_or,
                                       sampler=sampler, **kw)
    np.testing.assert_allclose(res["nhwc"]["loss"], l_or, rtol=RTOL)
    np.testing.assert_allclose(gn, g_or, rtol=RTOL, atol=2e-5 * np.abs(g_or).max())
    np.testing.assert_allclose(res["nhwc"]["protos"].cpu().numpy(), p_or, rtol=RTOL, atol=1e-6)


def test_channels_last_cache_modes_and_fallbacks():
    import css_b200
    from css_b200 import synth
    B2, C, h, w = 2, 7, 24, 20
    d = synth.student_batch(B2, C, h, w, seed=2, strategy="mix", block=4)
    protos = synth.warm_prototypes(C, seed=3).cuda()
    rep = cl(d["rep"])
    label, mask = d["label"].cuda(), d["mask"].cuda()
    kw = dict(num_queries=8, num_negatives=16, temp=0.5, strong_threshold=0.8, seed=4)

    def run(rep_in, prob):
        crit = css_b200.Contrast_Loss(**kw).cuda()
        loss = crit(rep_in, label, mask, prob, protos.clone())
        return loss.item(), crit.last

    prob = css_b200.ops.proto_softmax_sim(rep, protos, 0.5)
    base, last = run(rep, prob)
    assert last["rows_cache_mode"] == "same" and last["rows"].data_ptr() == rep.data_ptr()
    l2, last = run(rep, prob.clone())                        # nothing carried: norms-only pass over the map
    assert last["rows_cache_mode"] == "miss" and l2 == base
    other = cl(d["rep"] * 1.5 + 0.25)                        # other content, fresh tensor: the device check fails, norms are redone
    prob = css_b200.ops.proto_softmax_sim(rep, protos, 0.5)
    l3, last = run(other, prob)
    assert last["rows_cache_mode"] == "verify" and int(last["ws"].meta[css_b200._lib.META_ROWS_STALE].item()) == 1
    l4, last4 = run(other, prob.clone())
    assert l3 == l4
    # a bfloat16 channels-last map is not a fast-path input: it is converted to NCHW (correct, one extra copy)
    rep16 = rep.to(torch.bfloat16)
    assert not css_b200.ops.is_channels_last(rep16)
    p16 = css_b200.ops.proto_softmax_sim(rep16, protos, 0.5)
    p32 = css_b200.ops.proto_softmax_sim(rep16.float().contiguous(), protos, 0.5)
    assert torch.equal(p16, p32)
