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
