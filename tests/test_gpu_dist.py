"""World-size-2 run of the CUDA Contrast_Loss on ONE GPU (gloo backend moves the CUDA class-statistics tensor), checked
against the oracle fed what the reference's all_gather would have produced (loss.py:77,81,102): batch sharding, the
[C, D+1] all-reduce and the rank-local prototype update rule (loss.py:96-97)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    import css_b200
    from oracle import css_oracle as O
    g = load_golden("loss_mix_c21")
    C, Q, Nn = int(g["C"]), int(g["Q"]), int(g["Nn"])
    rep, label, mask, prob = g["s0_rep"], g["s0_label"].astype(np.float32), g["s0_mask"].astype(np.float32), g["s0_prob"]
    sl = slice(rank, rank + 1)
    kw = dict(num_queries=Q, num_negatives=Nn, temp=float(g["temp"]), strong_threshold=float(g["strong"]), alpha=float(g["alpha"]))
    crit = css_b200.Contrast_Loss(seed=100 + rank, **kw).cuda()
    protos = torch.from_numpy(g["s0_proto_in"].copy()).cuda()
    rep_t = torch.from_numpy(rep[sl].copy()).cuda().requires_grad_(True)
    loss = crit(rep_t, torch.from_numpy(label[sl].copy()).cuda(), torch.from_numpy(mask[sl].copy()).cuda(),
                torch.from_numpy(prob[sl].copy()).cuda(), protos)
    loss.backward()
    sel = crit.selection()
    a, n = crit.sample_indices(100 + rank, 0)
    slots = [k for k in range(sel["V"]) if sel["n_hard"][k] > 0] if sel["V"] > 1 else []
    sampler = O.RecordedDraws([a.cpu().numpy()[k] for k in slots], [n.cpu().numpy()[k].reshape(-1) for k in slots])
    p_or = g["s0_proto_in"].copy()
    l_or, g_or, info = O.contrast_loss(rep[sl], label[sl], mask[sl], prob[sl], p_or, sampler=sampler, rep_gather=rep,
                                       valid_gather=label * mask, **kw)
    out[rank] = dict(proto_err=float(np.abs(protos.cpu().numpy() - p_or).max()),
                     loss=float(loss.item()), loss_ref=float(l_or),
                     grad_err=float(np.abs(rep_t.grad.cpu().numpy() - g_or).max()), grad_max=float(np.abs(g_or).max()),
                     present=sel["present"], present_ref=info["present"], mode=crit.exchange_mode())
    dist.destroy_process_group()


def test_contrast_loss_world2_allreduce():
    port = 29900 + (os.getpid() % 90)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in (0, 1):
        r = out[rank]
        assert r["present"] == r["present_ref"]
        assert r["proto_err"] < 2e-5, r
        assert abs(r["loss"] - r["loss_ref"]) <= 1e-4 * abs(r["loss_ref"]), r
        assert r["grad_err"] <= 1e-4 * r["grad_max"] + 1e-7, r
    # the two ranks see different local class sets, so their prototypes legitimately differ (reference behaviour)
    assert out[0]["mode"] == out[1]["mode"] and out[0]["mode"] in ("peer", "nccl")


def _peer_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from css_b200.comm import PeerStatsReducer
    C, D = 21, 256

    def block(it, r):
        return torch.randn(C, D + 1, generator=torch.Generator().manual_seed(1000 * it + r))

    def expected(it):
        s = torch.zeros(C, D + 1)
        for r in range(world):                      # rank order, as the kernel sums
            s = s + block(it, r)
        return s

    red = PeerStatsReducer("cuda:0")
    res = dict(ok=red.ok, errs=[], graph_errs=[])
    if red.ok:
        for it in range(5):                         # odd and even epochs: both slot parities
            x = block(it, rank).cuda()
            red.allreduce(x, C, D)
            res["errs"].append(float((x.cpu() - expected(it)).abs().max()))
        static = torch.zeros(C, D + 1).cuda()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                red.allreduce(static, C, D)
        for it in range(10, 13):                    # the epoch lives on the device: replays keep working
            static.copy_(block(it, rank))
            g.replay()
            torch.cuda.synchronize()
            res["graph_errs"].append(float((static.cpu() - expected(it)).abs().max()))
        dist.barrier()
        red.close()
    out[rank] = res
    dist.destroy_process_group()


def test_peer_memory_allreduce_world2():
    """css_stats_allreduce between two processes (CUDA IPC; both on GPU 0 here, one per GPU in production): rank-ordered sum,
    bit-identical on both ranks, over several epochs and through CUDA-graph replays."""
    port = 29800 + (os.getpid() % 90)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0]["ok"] == out[1]["ok"]
    if not out[0]["ok"]:
        pytest.skip("CUDA IPC between the two test processes is not available on this box")
    for rank in (0, 1):
        assert out[rank]["errs"] == [0.0] * 5, out[rank]
        assert out[rank]["graph_errs"] == [0.0] * 3, out[rank]


def _cutmix_worker(rank, world, port, out, mode="cutmix"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    import random
    from css_b200 import aug
    g = load_golden(f"cut_{mode}_2_world2")          # recorded from the reference's generate_cut_gather_2 on two ranks
    seed = int(g["seed"])
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)

    def dev(k, dtype=None):
        t = torch.from_numpy(g[f"r{rank}_{k}"]).cuda()
        return t.to(dtype) if dtype is not None else t

    got = aug.generate_cut_gather_2(dev("image"), dev("label0", torch.int64), dev("conf0"), dev("conf1"), mode=mode)
    ok = all(np.array_equal(a.cpu().numpy(), g[f"r{rank}_{k}"].astype(a.cpu().numpy().dtype))
             for a, k in zip(got, ("out_image", "out_label0", "out_conf0", "out_conf1")))
    out[rank] = bool(ok) and got[1].dtype == torch.int64
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["cutmix", "classmix"])
def test_generate_cut_gather_world2_uses_rank0_partners(mode):
    """Two ranks: partners are rank 0's images; boxes (cutmix) and class permutations (classmix: only the other rank's counts of
    label values travel) are drawn for every gathered image in the reference's order -- outputs equal the reference's recorded ones."""
    port = 29700 + (os.getpid() % 90) + (100 if mode == "classmix" else 0)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_cutmix_worker, args=(2, port, out, mode), nprocs=2, join=True)
    assert out[0] and out[1]


def _timeout_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CSS_B200_COMM_TIMEOUT_S"] = "0.4" if rank == 0 else "600"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    import time
    from css_b200.comm import PeerStatsReducer
    C, D = 21, 256
    red = PeerStatsReducer("cuda:0")
    res = dict(ok=red.ok)
    if red.ok:
        mine = torch.full((C, D + 1), float(rank + 1)).cuda()
        x = mine.clone()
        if rank == 1:
            time.sleep(2.5)                          # a late rank (checkpoint write, loader respawn, ...)
        red.allreduce(x, C, D)
        torch.cuda.synchronize()
        res["timeouts"] = red.timeouts()
        res["finite"] = bool(torch.isfinite(x).all().item())
        if rank == 0:                                # gave up after 0.4 s: statistics stay LOCAL (never NaN), the event is counted,
            res["kept_local"] = bool(torch.equal(x, mine))      # and the next call raises instead of training on
            try:
                red.allreduce(x, C, D)
                res["raised"] = False
            except RuntimeError as e:
                res["raised"] = "timed out" in str(e)
        else:                                        # the late rank finds rank 0's block waiting for it: full sum
            res["summed"] = bool(torch.equal(x, torch.full((C, D + 1), 3.0).cuda()))
        dist.barrier()
        red.close()
    out[rank] = res
    dist.destroy_process_group()


def test_peer_exchange_timeout_is_counted_not_poisoned():
    """a rank that is late by more than the timeout must not turn the prototypes into NaN silently."""
    port = 29600 + (os.getpid() % 90)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_timeout_worker, args=(2, port, out), nprocs=2, join=True)
    if not out[0]["ok"]:
        pytest.skip("CUDA IPC between the two test processes is not available on this box")
    assert out[0]["timeouts"] == 1 and out[0]["kept_local"] and out[0]["finite"] and out[0]["raised"], dict(out[0])
    assert out[1]["timeouts"] == 0 and out[1]["summed"], dict(out[1])


def _sync_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    import css_b200
    g = load_golden("loss_mix_c21")
    Q, Nn = int(g["Q"]), int(g["Nn"])
    sl = slice(rank, rank + 1)
    res = {}
    for sync in (False, True):
        crit = css_b200.Contrast_Loss(num_queries=Q, num_negatives=Nn, temp=float(g["temp"]), strong_threshold=float(g["strong"]),
                                      alpha=float(g["alpha"]), seed=7 + rank, sync_prototypes=sync).cuda()
        protos = torch.from_numpy(g["s0_proto_in"].copy()).cuda()
        for _ in range(2):
            with torch.no_grad():
                crit(torch.from_numpy(g["s0_rep"][sl].copy()).cuda(), torch.from_numpy(g["s0_label"][sl].astype(np.float32)).cuda(),
                     torch.from_numpy(g["s0_mask"][sl].astype(np.float32)).cuda(), torch.from_numpy(g["s0_prob"][sl].copy()).cuda(), protos)
        res[sync] = protos.cpu().numpy()
        res[("present", sync)] = crit.selection()["present"]
    out[rank] = res
    dist.destroy_process_group()


def test_sync_prototypes_keeps_ranks_bit_identical():
    """sync_prototypes=True updates every globally present class on every rank from the summed statistics
    (no broadcast after the fact), so equal prototypes stay equal; the default keeps the reference's rank-local rule."""
    port = 29500 + (os.getpid() % 90)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sync_worker, args=(2, port, out), nprocs=2, join=True)
    assert np.array_equal(out[0][True], out[1][True])
    if out[0][("present", False)] != out[1][("present", False)]:        # different local class sets: the reference rule drifts
        assert not np.array_equal(out[0][False], out[1][False])
    # classes present on both ranks get the same update under either rule
    both = sorted(set(out[0][("present", False)]) & set(out[1][("present", False)]))
    assert both and np.array_equal(out[0][False][both], out[0][True][both])
