"""GPU parity of Contrast_Loss (selection, class statistics, prototype EMA, sampling, scoring + CE, backward) through the
C ABI, against the golden bundles recorded from the live reference (its own RNG draws fed back) and against the numpy
oracle on larger seeded inputs.  Selection lists / counts / index->pixel mapping: exactly equal.  Loss, gradient,
prototypes: rel 1e-4 (BASELINE.json north_star fp32 tolerance)."""
import numpy as np
import pytest
import torch

from oracle import css_oracle as O
from tests.helpers import load_golden, recorded, slot_major

pytestmark = pytest.mark.gpu

RTOL = 1e-4
LOSS_CASES = ["loss_ori_c5", "loss_mix_c21", "loss_nohard_c7", "loss_single_c4", "loss_cross_c19"]


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def scored_slots(sel):
    return [k for k in range(sel["V"]) if sel["n_hard"][k] > 0] if sel["V"] > 1 else []


def run_gpu(crit, rep, label, mask, prob, protos, indices=None):
    rep_t = dev(rep).requires_grad_(True)
    loss = crit(rep_t, dev(label), dev(mask), dev(prob), protos, _indices=indices)
    loss.backward()
    return loss, rep_t.grad


def check_selection(sel, info):
    assert sel["V"] == info["V"] and sel["present"] == info["present"] and sel["num_list"] == info["num_list"]
    for k in range(sel["V"]):
        assert np.array_equal(sel["valid_ids"][k], info["valid_ids"][k]), f"valid list of slot {k}"
        assert np.array_equal(sel["hard_ids"][k], info["hard_ids"][k]), f"hard list of slot {k}"


@pytest.mark.parametrize("name", LOSS_CASES)
def test_contrast_loss_against_reference_golden(name):
    """Reference's recorded indices fed to the CUDA path: loss / grad / prototypes vs the reference's own outputs."""
    import css_b200
    g = load_golden(name)
    C, Q, Nn = int(g["C"]), int(g["Q"]), int(g["Nn"])
    kw = dict(num_queries=Q, num_negatives=Nn, temp=float(g["temp"]), strong_threshold=float(g["strong"]), alpha=float(g["alpha"]))
    crit = css_b200.Contrast_Loss(**kw).cuda()
    for s in range(int(g["steps"])):
        rep, label, mask, prob = g[f"s{s}_rep"], g[f"s{s}_label"].astype(np.float32), g[f"s{s}_mask"].astype(np.float32), g[f"s{s}_prob"]
        a_rec, n_rec = recorded(g, s)
        # selection does not depend on the draws: run once without gradient to learn the scored slots
        protos = dev(g[f"s{s}_proto_in"])
        with torch.no_grad():
            crit(dev(rep), dev(label), dev(mask), dev(prob), protos.clone())
        sel = crit.selection()
        p_or = g[f"s{s}_proto_in"].copy()
        l_or, g_or, info = O.contrast_loss(rep, label, mask, prob, p_or, sampler=O.RecordedDraws(a_rec, n_rec), **kw)
        check_selection(sel, info)
        slots = scored_slots(sel)
        assert slots == info["scored"] and len(slots) == int(g[f"s{s}_n_scored"])
        a, n = slot_major(a_rec, n_rec, slots, C, Q, Nn)
        loss, grad = run_gpu(crit, rep, label, mask, prob, protos, (dev(a), dev(n)))
        assert loss.dim() == 0 and loss.dtype == torch.float32
        np.testing.assert_allclose(protos.cpu().numpy(), g[f"s{s}_proto_out"], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(loss.item(), g[f"s{s}_loss"], rtol=RTOL, atol=1e-6)
        gref = g[f"s{s}_grad"]
        gg = grad.cpu().numpy()
        assert np.array_equal(gg != 0, gref != 0), "gradient support differs from the reference"
        np.testing.assert_allclose(gg, gref, rtol=RTOL, atol=1e-7)
        # anchor pixels are exactly the reference's
        apx = crit.last["anchor_px"].cpu().numpy().reshape(C, Q)
        for k, px in zip(info["scored"], info["anchor_pixels"]):
            assert np.array_equal(apx[k], px)
        if len(slots) == 0:
            assert loss.item() == 0.0 and not gg.any()


@pytest.mark.parametrize("B2,C,h,w,Q,Nn,strategy,strong", [(2, 21, 81, 81, 256, 512, "ori", 0.97),     # BASELINE config 1
                                                           (2, 19, 33, 47, 64, 100, "mix", 0.5),
                                                           (3, 32, 20, 20, 7, 3, "ori", 0.9)])
def test_contrast_loss_against_oracle_device_draws(B2, C, h, w, Q, Nn, strategy, strong):
    """Draws come from the device sampler (css_sample), are fed to the oracle, and the on-the-fly path must reproduce the
    fed path bit for bit."""
    import css_b200
    from css_b200 import synth
    d = synth.student_batch(B2, C, h, w, seed=77 + C, strategy=strategy, block=8 if h > 40 else 3)
    protos0 = synth.warm_prototypes(C, seed=5, zero_rows=(1,))
    if strategy == "mix":
        d["prob"] = torch.from_numpy(O.proto_softmax_sim(d["rep"].numpy(), protos0.numpy(), 0.5))
    kw = dict(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=strong, alpha=0.99)
    crit = css_b200.Contrast_Loss(seed=1234, **kw).cuda()
    rep, label, mask, prob = (d[k].numpy() for k in ("rep", "label", "mask", "prob"))
    protos = protos0.clone().cuda()
    loss_fly, grad_fly = run_gpu(crit, rep, label, mask, prob, protos)          # on-the-fly draws, offset 0
    sel = crit.selection()
    a, n = crit.sample_indices(1234, 0)
    slots = scored_slots(sel)
    a_np, n_np = a.cpu().numpy(), n.cpu().numpy()
    for k in range(C):
        if k in slots:
            assert a_np[k].min() >= 0 and a_np[k].max() < sel["n_hard"][k]
            tot = sum(sel["num_list"]) - sel["num_list"][k]
            assert n_np[k].min() >= 0 and n_np[k].max() < tot
        else:
            assert (a_np[k] == -1).all() and (n_np[k] == -1).all()
    p_or = protos0.numpy().copy()
    sampler = O.RecordedDraws([a_np[k] for k in slots], [n_np[k].reshape(-1) for k in slots])
    l_or, g_or, info = O.contrast_loss(rep, label, mask, prob, p_or, sampler=sampler, **kw)
    check_selection(sel, info)
    assert info["scored"] == slots
    np.testing.assert_allclose(protos.cpu().numpy(), p_or, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(loss_fly.item(), l_or, rtol=RTOL)
    gg = grad_fly.cpu().numpy()
    assert np.array_equal(gg != 0, g_or != 0)
    np.testing.assert_allclose(gg, g_or, rtol=RTOL, atol=1e-7)
    # fed path == on-the-fly path, bit for bit (loss) and to rounding of the atomic accumulation order (grad)
    protos2 = protos0.clone().cuda()
    loss_fed, grad_fed = run_gpu(crit, rep, label, mask, prob, protos2, (a, n))
    assert loss_fed.item() == loss_fly.item()
    assert torch.equal(protos2, protos)
    torch.testing.assert_close(grad_fed, grad_fly, rtol=1e-6, atol=1e-9)
    # a different offset gives different draws
    a2, _ = crit.sample_indices(1234, 1)
    assert not torch.equal(a2, a)


def test_multi_hot_labels_and_empty_batch():
    """label need not be one-hot for the reference (a pixel may be valid for several classes); an all-masked batch
    gives V = 0, loss exactly 0, dense zero gradient, prototypes untouched."""
    import css_b200
    g = torch.Generator().manual_seed(3)
    B2, C, h, w, Q, Nn = 2, 6, 12, 10, 8, 16
    rep = torch.randn(B2, 256, h, w, generator=g).numpy()
    label = (torch.rand(B2, C, h, w, generator=g) < 0.3).float().numpy()
    mask = (torch.rand(B2, 1, h, w, generator=g) < 0.7).float().numpy()
    prob = torch.rand(B2, C, h, w, generator=g).numpy()
    kw = dict(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.6, alpha=0.99)
    crit = css_b200.Contrast_Loss(seed=7, **kw).cuda()
    protos = torch.zeros(C, 256).cuda()
    loss, grad = run_gpu(crit, rep, label, mask, prob, protos)
    sel = crit.selection()
    a, n = crit.sample_indices(7, 0)
    slots = scored_slots(sel)
    p_or = np.zeros((C, 256), np.float32)
    sampler = O.RecordedDraws([a.cpu().numpy()[k] for k in slots], [n.cpu().numpy()[k].reshape(-1) for k in slots])
    l_or, g_or, info = O.contrast_loss(rep, label, mask, prob, p_or, sampler=sampler, **kw)
    check_selection(sel, info)
    np.testing.assert_allclose(loss.item(), l_or, rtol=RTOL)
    np.testing.assert_allclose(grad.cpu().numpy(), g_or, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(protos.cpu().numpy(), p_or, rtol=RTOL, atol=1e-6)
    # empty
    protos = torch.ones(C, 256).cuda()
    loss, grad = run_gpu(crit, rep, label, np.zeros_like(mask), prob, protos)
    assert loss.item() == 0.0 and not grad.any().item() and grad.shape == (B2, 256, h, w)
    assert crit.selection()["V"] == 0 and torch.equal(protos, torch.ones(C, 256).cuda())


def uniform_pvalue(idx, n, max_bins=8):
    """chi-square p-value of `idx` being uniform on [0, n), with coarse bins whose expected mass follows their width."""
    from scipy import stats
    bins = min(max_bins, n)
    which = (np.arange(n, dtype=np.int64) * bins) // n
    width = np.bincount(which, minlength=bins).astype(np.float64)
    cnt = np.bincount(which[np.asarray(idx, dtype=np.int64)], minlength=bins)
    return stats.chisquare(cnt, width / width.sum() * cnt.sum()).pvalue


def test_sampler_distributions():
    """chi-square tests of the device Philox sampler: anchor uniformity, class histogram vs
    softmax(cos(P_k,P_j)/temp), uniformity inside a class."""
    import css_b200
    from css_b200 import synth
    from scipy import stats
    B2, C, h, w, Q, Nn = 2, 5, 16, 16, 512, 256
    d = synth.student_batch(B2, C, h, w, seed=11, block=4)
    crit = css_b200.Contrast_Loss(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.97, seed=99).cuda()
    protos = synth.warm_prototypes(C, seed=2).cuda()
    with torch.no_grad():
        crit(d["rep"].cuda(), d["label"].cuda(), d["mask"].cuda(), d["prob"].cuda(), protos)
    sel = crit.selection()
    assert sel["V"] == C
    a, n = crit.sample_indices(99, 0)
    a, n = a.cpu().numpy(), n.cpu().numpy()
    proto_rep = protos.cpu().numpy()
    for k in range(C):
        if sel["n_hard"][k] == 0:
            continue
        # (a) anchors uniform over the hard list (coarse bins so that expected counts are large)
        assert uniform_pvalue(a[k], sel["n_hard"][k]) > 1e-4
        # (b) class histogram of the negatives vs proto_prob
        others = sel["num_list"][k + 1:] + sel["num_list"][:k]
        edges = np.concatenate([[0], np.cumsum(others)])
        seg = np.searchsorted(edges, n[k].reshape(-1), side="right") - 1
        obs = np.bincount(seg, minlength=C - 1)
        p = O.proto_class_prob(proto_rep[sel["present"]], k, 0.5).astype(np.float64)
        assert stats.chisquare(obs, p / p.sum() * obs.sum()).pvalue > 1e-4
        # (c) uniform inside the largest segment
        j = int(np.argmax(others))
        inside = n[k].reshape(-1)[seg == j] - edges[j]
        assert uniform_pvalue(inside, others[j]) > 1e-4


def test_full_size_properties_voc_batch():
    """BASELINE config 2 shape (B2=16, C=21, 81x81, Q=256, Nn=512): size-independent properties instead of the oracle."""
    import css_b200
    from css_b200 import synth
    B2, C, h, w, Q, Nn = 16, 21, 81, 81, 256, 512
    d = synth.student_batch(B2, C, h, w, seed=3407)
    crit = css_b200.Contrast_Loss(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.97, seed=5).cuda()
    rep = d["rep"].cuda().requires_grad_(True)
    label, mask, prob = d["label"].cuda(), d["mask"].cuda(), d["prob"].cuda()
    protos = torch.zeros(C, 256).cuda()
    loss = crit(rep, label, mask, prob, protos)
    loss.backward()
    sel = crit.selection()
    valid = (label * mask) != 0
    # counts / sortedness / membership of the compaction
    assert sel["num_list"] == [int(valid[:, c].sum()) for c in sel["present"]]
    for k, c in enumerate(sel["present"]):
        ids = sel["valid_ids"][k]
        assert np.all(np.diff(ids) > 0)
        assert np.array_equal(ids, torch.nonzero(valid[:, c].reshape(-1)).reshape(-1).cpu().numpy())
        hard = valid[:, c] & (prob[:, c] < 0.97)
        assert np.array_equal(sel["hard_ids"][k], torch.nonzero(hard.reshape(-1)).reshape(-1).cpu().numpy())
    # first-touch prototypes are the class means of the raw features
    x = d["rep"].permute(0, 2, 3, 1).reshape(-1, 256)
    for k, c in enumerate(sel["present"]):
        np.testing.assert_allclose(protos[c].cpu().numpy(), x[sel["valid_ids"][k]].double().mean(0).numpy(), rtol=1e-4, atol=1e-5)
    # gradient support == anchor pixels; loss positive and below log(1+Nn) + 2/temp
    apx = crit.last["anchor_px"].cpu().numpy()
    apx = np.unique(apx[apx >= 0])
    gpix = torch.nonzero(rep.grad.abs().sum(1).reshape(-1)).reshape(-1).cpu().numpy()
    assert set(gpix).issubset(set(apx)) and len(gpix) > 0.9 * len(apx)
    assert 0 < loss.item() < np.log(1 + Nn) + 4
    # cosine scoring is invariant to a positive rescaling of rep; prototypes (raw means) scale linearly
    crit2 = css_b200.Contrast_Loss(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.97, seed=5).cuda()
    protos2 = torch.zeros(C, 256).cuda()
    loss2 = crit2((d["rep"] * 4).cuda(), label, mask, prob, protos2)
    np.testing.assert_allclose(loss2.item(), loss.item(), rtol=1e-5)
    torch.testing.assert_close(protos2, protos * 4, rtol=1e-5, atol=1e-6)
    # same seed/offset -> same loss bit for bit (deterministic reductions everywhere on the forward path)
    crit3 = css_b200.Contrast_Loss(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.97, seed=5).cuda()
    protos3 = torch.zeros(C, 256).cuda()
    assert crit3(d["rep"].cuda(), label, mask, prob, protos3).item() == loss.item()
    assert torch.equal(protos3, protos)
    # EMA branch on a second step with a DIFFERENT batch: p <- alpha p + (1-alpha) mean(second batch) (loss.py:102,108)
    before = protos.clone()
    d2 = synth.student_batch(B2, C, h, w, seed=99)
    crit(d2["rep"].cuda(), d2["label"].cuda(), d2["mask"].cuda(), d2["prob"].cuda(), protos)
    x2 = d2["rep"].permute(0, 2, 3, 1).reshape(-1, 256).double()
    valid2 = ((d2["label"] * d2["mask"]) != 0).permute(0, 2, 3, 1).reshape(-1, C)
    assert not torch.equal(before, protos)
    for c in range(C):
        want = 0.99 * before[c].double().cpu() + 0.01 * x2[valid2[:, c]].mean(0) if valid2[:, c].any() else before[c].double().cpu()
        np.testing.assert_allclose(protos[c].cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-6)


def test_model_shells_with_stub_network():
    """Model_mix / Model_cross / Model_ori_pseudo shells (stub network, pass-through aug) reproduce the reference tuples
    recorded in the stage-1/2 golden bundles."""
    import css_b200
    from css_b200 import models
    from tests.helpers import top2_margin

    class Stub(torch.nn.Module):
        def __init__(self, outs):
            super().__init__()
            self.outs, self.i = outs, 0

        def forward(self, x):
            o = self.outs[self.i % len(self.outs)]
            self.i += 1
            return o

    saved = dict(vars(models.hooks))
    try:
        models.hooks.network_factory = lambda enc, **kw: torch.nn.Conv2d(1, 1, 1)
        models.hooks.batch_transform = lambda a, b, c, **k: (a, b, c)
        models.hooks.batch_transform_2 = lambda a, b, c, d, **k: (a, b, c, d)
        models.hooks.batch_transform_3 = lambda a, b, c, d, e, **k: (a, b, c, d, e)
        models.hooks.generate_cut_gather = lambda a, b, c, mode=None: (a, b, c)
        models.hooks.generate_cut_gather_2 = lambda a, b, c, d, mode=None: (a, b, c, d)
        models.hooks.generate_cut_gather_3 = lambda a, b, c, d, e, mode=None: (a, b, c, d, e)
        for name, cls in (("stage12_mix_c21", models.Model_mix), ("stage12_cross_c19", models.Model_cross)):
            g = load_golden(name)
            B, C, H, W, temp = int(g["B"]), int(g["C"]), int(g["H"]), int(g["W"]), float(g["temp"])
            cfg = {"Dataset": {"crop_size": (H, W), "scale_size": (1.0, 1.0), "mix_mode": "none"}}
            m = cls(None, num_classes=C, output_dim=256, config=cfg, temp=temp).cuda()
            rep_all, pred_all = dev(g["rep_all"]), dev(g["pred_all"])
            m.ema_model = Stub([(dev(g["pred_u"]), dev(g["rep_u"]))])
            m.model = Stub([(pred_all[:B], rep_all[:B]), (pred_all[B:], rep_all[B:])])
            img = torch.zeros(B, 3, H, W).cuda()
            r = m(img, img, dev(g["prototypes"]))
            assert len(r) == (7 if cls is models.Model_mix else 8)
            np.testing.assert_allclose(r[-1].cpu().numpy(), g["prob_all"], atol=1e-6)
            assert torch.equal(r[-2], rep_all)
            assert r[0].shape == (B, C, H, W) and r[1].shape == (B, C, H, W)
            if cls is models.Model_mix:
                assert r[2].dtype == torch.float32
                np.testing.assert_allclose(r[3].cpu().numpy(), g["conf_cls"], atol=2e-6)
                np.testing.assert_allclose(r[4].cpu().numpy(), g["conf_rep"], atol=2e-6)
                assert (r[2].cpu().numpy() != g["fused"]).mean() < 1e-3
            else:
                assert r[2].dtype == torch.int64 and r[3].dtype == torch.int64
                assert (r[2].cpu().numpy() != g["label_cls"]).mean() < 1e-3
                assert (r[3].cpu().numpy() != g["label_rep"]).mean() < 1e-3
        m.step = 0
        m.ema_update()
        assert m.step == 1
    finally:
        for k, v in saved.items():
            setattr(models.hooks, k, v)


def test_bf16_rep_matches_fp32_path_on_widened_values():
    """bf16 representation maps (north_star extension, tolerance rel 1e-2): the kernels widen bf16 exactly and compute in
    fp32, so the result must equal the fp32 path on rep.float() -- far inside the stated 1e-2."""
    import css_b200
    from css_b200 import synth
    B2, C, h, w, Q, Nn = 2, 21, 33, 31, 32, 64
    d = synth.student_batch(B2, C, h, w, seed=9, strategy="mix", block=4)
    rep16 = d["rep"].to(torch.bfloat16).cuda()
    protos0 = synth.warm_prototypes(C, seed=4)
    prob16 = css_b200.ops.proto_softmax_sim(rep16, protos0.cuda(), 0.5)
    prob32 = css_b200.ops.proto_softmax_sim(rep16.float(), protos0.cuda(), 0.5)
    assert torch.equal(prob16, prob32)
    np.testing.assert_allclose(prob16.cpu().numpy(), O.proto_softmax_sim(rep16.float().cpu().numpy(), protos0.numpy(), 0.5), atol=1e-6)
    outs = []
    for rep in (rep16, rep16.float()):
        crit = css_b200.Contrast_Loss(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.8, seed=3).cuda()
        protos = protos0.clone().cuda()
        r = rep.clone().requires_grad_(True)
        loss = crit(r, d["label"].cuda(), d["mask"].cuda(), prob16, protos)
        loss.backward()
        assert r.grad.dtype == rep.dtype
        outs.append((loss.item(), r.grad.float(), protos))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][2], outs[1][2])
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=1e-2, atol=1e-6)


def test_cuda_graph_replays_draw_fresh_samples():
    """The whole forward+backward is capturable (no host sync, caller-owned buffers) and the Philox offset lives on the
    device, so every replay of the captured graph draws new anchors / negatives, identical to the eager call at that offset."""
    import css_b200
    from css_b200 import synth
    B2, C, h, w, Q, Nn = 2, 7, 24, 24, 16, 32
    d = synth.student_batch(B2, C, h, w, seed=21, block=4)
    rep, label, mask, prob = d["rep"].cuda(), d["label"].cuda(), d["mask"].cuda(), d["prob"].cuda()
    kw = dict(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.97)
    crit = css_b200.Contrast_Loss(seed=77, **kw).cuda()
    protos = synth.warm_prototypes(C, seed=1).cuda()

    def step():
        r = rep.detach().requires_grad_(True)
        loss = crit(r, label, mask, prob, protos)
        (g,) = torch.autograd.grad(loss, r)
        return loss, g

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()                                   # offsets 0, 1
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_g, grad_g = step()                      # capture does not execute: the counter stays at 2
    assert crit.draw_offset() == 2
    seen = []
    for _ in range(3):
        graph.replay()
        seen.append((loss_g.item(), grad_g.clone()))
    assert crit.draw_offset() == 5
    assert len({s[0] for s in seen}) == 3, "graph replays repeated the same draws"
    # eager calls at offsets 2, 3, 4 with identically evolving prototypes reproduce the replays bit for bit
    crit2 = css_b200.Contrast_Loss(seed=77, **kw).cuda()
    crit2.set_sampler(77, step=0)
    protos2 = synth.warm_prototypes(C, seed=1).cuda()
    losses2 = []
    for i in range(5):
        r = rep.detach().requires_grad_(True)
        l2 = crit2(r, label, mask, prob, protos2)
        losses2.append(l2.item())
    assert losses2[2:] == [s[0] for s in seen]
    assert torch.equal(protos2, protos)


@pytest.mark.parametrize("temp", [0.02, 0.1])
def test_small_temperatures_online_max_path(temp):
    """temp = 0.02 takes the online-max softmax path of the scorer (fixed-reference softmax is only used while
    2^(-2 log2(e)/temp) is far from fp32 underflow); both must match the oracle."""
    import css_b200
    from css_b200 import synth
    B2, C, h, w, Q, Nn = 2, 9, 20, 18, 16, 40
    d = synth.student_batch(B2, C, h, w, seed=31, block=3)
    kw = dict(num_queries=Q, num_negatives=Nn, temp=temp, strong_threshold=0.97, alpha=0.99)
    crit = css_b200.Contrast_Loss(seed=5, **kw).cuda()
    rep, label, mask, prob = (d[k].numpy() for k in ("rep", "label", "mask", "prob"))
    protos0 = synth.warm_prototypes(C, seed=8)
    protos = protos0.clone().cuda()
    loss, grad = run_gpu(crit, rep, label, mask, prob, protos)
    sel = crit.selection()
    a, n = crit.sample_indices(5, 0)
    slots = scored_slots(sel)
    sampler = O.RecordedDraws([a.cpu().numpy()[k] for k in slots], [n.cpu().numpy()[k].reshape(-1) for k in slots])
    p_or = protos0.numpy().copy()
    l_or, g_or, info = O.contrast_loss(rep, label, mask, prob, p_or, sampler=sampler, **kw)
    np.testing.assert_allclose(loss.item(), l_or, rtol=RTOL)
    np.testing.assert_allclose(grad.cpu().numpy(), g_or, rtol=RTOL, atol=1e-6 * np.abs(g_or).max())


@pytest.mark.parametrize("B2,C,h,w,Q,Nn", [(1, 2, 1, 3, 1, 1), (1, 32, 3, 2, 3, 5), (2, 3, 1, 1, 2, 2), (1, 4, 5, 7, 130, 1), (3, 21, 4, 4, 1, 700)])
def test_tiny_and_ragged_shapes(B2, C, h, w, Q, Nn):
    """Degenerate sizes: single-row maps, more classes than pixels, one query, one negative, Q not a multiple of 4,
    Nn spanning several candidate batches.  Same checks as the large cases (selection exact, loss/grad rel 1e-4)."""
    import css_b200
    g = torch.Generator().manual_seed(B2 * 1000 + C * 10 + Q)
    rep = torch.randn(B2, 256, h, w, generator=g).numpy()
    cls = torch.randint(0, C, (B2, h, w), generator=g)
    label = torch.nn.functional.one_hot(cls, C).permute(0, 3, 1, 2).float().numpy()
    mask = (torch.rand(B2, 1, h, w, generator=g) < 0.9).float().numpy()
    prob = torch.rand(B2, C, h, w, generator=g).numpy()
    kw = dict(num_queries=Q, num_negatives=Nn, temp=0.5, strong_threshold=0.7, alpha=0.99)
    crit = css_b200.Contrast_Loss(seed=3, **kw).cuda()
    protos = torch.zeros(C, 256).cuda()
    version = protos._version
    loss, grad = run_gpu(crit, rep, label, mask, prob, protos)
    assert protos._version > version                # the in-place prototype update is visible to autograd's version counter
    sel = crit.selection()
    a, n = crit.sample_indices(3, 0)
    slots = scored_slots(sel)
    sampler = O.RecordedDraws([a.cpu().numpy()[k] for k in slots], [n.cpu().numpy()[k].reshape(-1) for k in slots])
    p_or = np.zeros((C, 256), np.float32)
    l_or, g_or, info = O.contrast_loss(rep, label, mask, prob, p_or, sampler=sampler, **kw)
    check_selection(sel, info)
    assert info["scored"] == slots
    np.testing.assert_allclose(loss.item(), l_or, rtol=RTOL, atol=1e-7)
    # d loss / d anchor = (S - p_hat) k - t a_hat is a difference of nearly equal vectors when hundreds of negatives are drawn
    # from a few dozen pixels: elements far below the row's scale carry fp32 summation-order noise of ~1e-5 of that scale
    np.testing.assert_allclose(grad.cpu().numpy(), g_or, rtol=RTOL, atol=2e-5 * max(np.abs(g_or).max(), 1e-30))
    np.testing.assert_allclose(protos.cpu().numpy(), p_or, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("Nn", [1, 5, 50, 100, 512, 700])
@pytest.mark.parametrize("temp", [0.5, 0.02])
def test_scorer_paths_agree(Nn, temp):
    """css_set_scorer_path: the shared-memory ring kernel and the register / bulk-copy hybrid draw the same candidates as the
    register kernel and agree with it to rounding (another summation order), on full and ragged candidate batches, with the
    draws made on the fly and fed, with the fixed-reference and the online-max softmax (temp 0.02)."""
    import css_b200
    from css_b200 import _lib, synth
    lib = _lib.load()
    B2, C, h, w, Q = 2, 7, 19, 23, 12
    d = synth.student_batch(B2, C, h, w, seed=11 + Nn, block=3)
    kw = dict(num_queries=Q, num_negatives=Nn, temp=temp, strong_threshold=0.97, alpha=0.99)
    rep, label, mask, prob = (d[k].numpy() for k in ("rep", "label", "mask", "prob"))
    protos0 = synth.warm_prototypes(C, seed=8)
    out = {}
    try:
        for path in (0, 1, 2):
            lib.css_set_scorer_path(path)
            crit = css_b200.Contrast_Loss(seed=9, **kw).cuda()
            loss, grad = run_gpu(crit, rep, label, mask, prob, protos0.clone().cuda())
            a, n = crit.sample_indices(9, 0)
            loss_fed, grad_fed = run_gpu(crit, rep, label, mask, prob, protos0.clone().cuda(), (a, n))
            assert loss_fed.item() == loss.item()
            torch.testing.assert_close(grad_fed, grad, rtol=1e-6, atol=1e-9)
            out[path] = (loss.item(), grad.cpu().numpy(), crit.last["anchor_px"].cpu().numpy())
    finally:
        lib.css_set_scorer_path(-1)
    l0, g0, px0 = out[0]
    assert np.abs(g0).max() > 0
    for path in (1, 2):
        l, g, px = out[path]
        assert np.array_equal(px, px0)
        np.testing.assert_allclose(l, l0, rtol=2e-6)
        np.testing.assert_allclose(g, g0, rtol=1e-5, atol=2e-6 * np.abs(g0).max())
