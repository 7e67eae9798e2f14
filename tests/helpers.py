"""Shared helpers of the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAU = 1e-5     # near-tie margin


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def recorded(g, s):
    n = int(g[f"s{s}_n_scored"])
    return [g[f"s{s}_anchor_idx_{k}"] for k in range(n)], [g[f"s{s}_neg_idx_{k}"] for k in range(n)]


def slot_major(anchor_list, neg_list, scored_slots, C, Q, Nn):
    """Recorded per-scored-class draws -> the slot-major int32 arrays css_score_ce is fed ([C,Q], [C,Q,Nn])."""
    a = np.full((C, Q), -1, dtype=np.int32)
    n = np.full((C, Q, Nn), -1, dtype=np.int32)
    for k, ai, ni in zip(scored_slots, anchor_list, neg_list):
        a[k] = np.asarray(ai, dtype=np.int32)
        n[k] = np.asarray(ni, dtype=np.int32).reshape(Q, Nn)
    return a, n


def top2_margin(p):
    s = np.sort(p, axis=1)
    return s[:, -1] - s[:, -2]


def assert_labels_match(label, label_ref, margin, what):
    """Labels identical wherever the reference's own top-2 margin >= TAU; mismatches below TAU are counted."""
    bad = label != label_ref
    assert (margin[bad] < TAU).all(), f"{what}: {int((margin[bad] >= TAU).sum())} label mismatches away from near-ties"
    return int(bad.sum())
