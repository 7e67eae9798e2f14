"""World-size-2 gloo test (CPU) of the one exchange step of the path: the all-reduce of per-class sums | counts that
replaces the reference's all_gather (loss.py:77,81,102), and the rank-local update rule (loss.py:96-97)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import load_golden


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import css_oracle as O
    from css_b200.loss import allreduce_class_stats
    g = load_golden("loss_mix_c21")
    rep, label, mask = g["s0_rep"], g["s0_label"].astype(np.float32), g["s0_mask"].astype(np.float32)
    sl = slice(rank, rank + 1)                                 # batch sharding: one image per rank
    sums, cnt = O.class_statistics(rep[sl], label[sl], mask[sl])
    stats = torch.from_numpy(np.concatenate([sums, cnt[:, None]], axis=1))
    local_cnt = cnt.copy()
    allreduce_class_stats(stats)                               # the product's host-side exchange step
    stats = stats.numpy()
    mean = stats[:, :-1] / np.maximum(stats[:, -1:], 1)
    # reference semantics on this rank: all_gather'ed rep / valid, local presence rule
    protos = g["s0_proto_in"].copy()
    O.contrast_loss(rep[sl], label[sl], mask[sl], g["s0_prob"][sl], protos, num_queries=4, num_negatives=8,
                    sampler=_NoDraws(), rep_gather=rep, valid_gather=label * mask, want_grad=False)
    mine = g["s0_proto_in"].copy()
    for c in range(mine.shape[0]):
        if local_cnt[c] > 0:
            mine[c] = mean[c] if mine[c].sum() == 0 else np.float32(0.99) * mine[c] + np.float32(1 - 0.99) * mean[c]
    out[rank] = (float(np.abs(mine - protos).max()), int((local_cnt > 0).sum()), float(stats[:, -1].sum()))
    dist.destroy_process_group()


class _NoDraws:
    def anchors(self, n_hard, Q):
        return np.zeros(Q, dtype=np.int64)

    def negatives(self, proto_prob, Q, Nn, negative_num_list):
        return np.zeros(Q * Nn, dtype=np.int64)


def test_allreduce_replaces_allgather_world2():
    port = 29500 + (os.getpid() % 400)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g = load_golden("loss_mix_c21")
    total_valid = float((g["s0_label"].astype(np.float32) * g["s0_mask"].astype(np.float32) != 0).sum())
    for rank in (0, 1):
        err, n_local, tot = out[rank]
        assert err < 1e-5, f"rank {rank}: prototypes differ from the all_gather semantics by {err}"
        assert n_local > 0 and tot == total_valid


def _cut_worker(rank, world, port, out, mode):
    """Host side of css_b200.aug.generate_cut_gather_2 on two ranks with CPU tensors: the device launch (aug.cut_mix, CUDA only)
    is replaced by a recorder, everything before it -- who draws what from which generator, what is gathered, where the partners
    come from -- runs as shipped."""
    import random
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from css_b200 import aug
    g = load_golden(f"cut_{mode}_2_world2")          # recorded from the reference's generate_cut_gather_2 on two ranks
    seed = int(g["seed"])
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    seen = {}

    def recorder(image, labels, confs, mode_, boxes=None, class_sets=None, partner=None):
        seen.update(mode=mode_, boxes=boxes, class_sets=class_sets, partner=partner)
        return image, labels, confs

    real, aug.cut_mix = aug.cut_mix, recorder
    try:
        t = lambda k, dt=None: torch.from_numpy(g[f"r{rank}_{k}"]).to(dt) if dt else torch.from_numpy(g[f"r{rank}_{k}"])   # noqa: E731
        aug.generate_cut_gather_2(t("image"), t("label0", torch.int64), t("conf0"), t("conf1"), mode=mode)
    finally:
        aug.cut_mix = real
    ok = seen["mode"] == mode and seen["partner"] is not None
    p_img, p_lab, p_conf = seen["partner"]
    ok &= np.array_equal(p_img.numpy(), g["r0_image"]) and np.array_equal(p_lab[0].numpy(), g["r0_label0"].astype(np.int64))
    ok &= np.array_equal(p_conf[0].numpy(), g["r0_conf0"]) and np.array_equal(p_conf[1].numpy(), g["r0_conf1"])
    if mode == "classmix":
        rec = [[int(v) for v in row if v != -100] for row in g[f"r{rank}_class_sets"]]
        ok &= [list(s) for s in seen["class_sets"]] == rec
    else:
        ok &= np.array_equal(np.asarray(seen["boxes"]), g[f"r{rank}_boxes"])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_cut_gather_world2_host_logic_cutmix_and_classmix():
    """N > 1 host logic of the augmentation hand-off on CPU (gloo, world size 2): partners are broadcast from rank 0, every rank
    draws boxes for all gathered images (CutMix), and for ClassMix only the per-image counts of label values are gathered while
    the permutation draws stay in the reference's order -- checked against what two ranks of the reference recorded."""
    for i, mode in enumerate(("cutmix", "classmix")):
        port = 29100 + (os.getpid() % 300) + 400 * i
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_cut_worker, args=(2, port, out, mode), nprocs=2, join=True)
        assert out[0] and out[1], f"{mode}: {dict(out)}"
