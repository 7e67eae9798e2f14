"""Harness that imports the UNMODIFIED reference (WangChangqi98/CSS).

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py in the build container to run
the live reference on CPU and record its outputs + RNG draws as golden fixtures, and by
oracle/ref_bench.py for bench.py's reference timings.  Nothing in ``css_b200/`` imports it.

Where the reference is looked for: $CSS_REFERENCE_ROOT, then ``baseline/_ref`` (a git-ignored copy
of the reference's Python tree that ``__graft_entry__.build()`` makes where /root/reference exists;
it travels to the GPU box with the snapshot), then /root/reference.

Shims (none of them edits the reference):
  1. single-process gloo group  -- ``concat_all_gather`` is unconditional
     (generalframeworks/loss/loss.py:77, networks/ddp_model.py:241-250)
  2. ``Tensor.cuda`` -> identity on GPU-less hosts (loss.py:147 calls ``.cuda()``)
  3. the three top-level scripts are never imported (``shutup`` is not installed).
"""
import os
import sys
import contextlib

import numpy as np
import torch
import torch.distributed as dist

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("CSS_REFERENCE_ROOT"), os.path.join(os.path.dirname(_HERE), "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "generalframeworks")):
            return cand
    return os.environ.get("CSS_REFERENCE_ROOT", "/root/reference")


REF_ROOT = _find_root()


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "generalframeworks"))


@contextlib.contextmanager
def cpu_cuda_shim():
    """Shim (2) as a scope: while the reference runs on CPU tensors on a box that HAS a GPU, loss.py:147's `.cuda()` must stay
    a no-op, and must be the real thing again afterwards."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


_mods = {}


def load_reference(device="cpu"):
    """Returns (loss_module, ddp_model_module, utils_module) of the reference."""
    if not dist.is_initialized():                  # concat_all_gather is unconditional: a world-1 group for CPU and CUDA tensors
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29571")
        dist.init_process_group("cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo", rank=0, world_size=1)
    if _mods:
        return _mods["L"], _mods["M"], _mods["U"]
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import generalframeworks.loss.loss as L
    import generalframeworks.networks.ddp_model as M
    import generalframeworks.utils as U
    _mods.update(L=L, M=M, U=U)
    return L, M, U


class DrawRecorder:
    """Records the three RNG streams Contrast_Loss.forward consumes, per present class:
    torch.randint (anchors, loss.py:127), Categorical.sample (loss.py:137) and
    negative_index_sampler (loss.py:140, 410-418)."""

    def __init__(self):
        self.anchor_idx = []   # list of int64 [Q]
        self.samp_class = []   # list of int64 [Q, Nn]
        self.neg_idx = []      # list of int64 [Q*Nn]

    @contextlib.contextmanager
    def recording(self):
        L, _, _ = load_reference()
        orig_randint = torch.randint
        orig_sampler = L.negative_index_sampler
        Cat = torch.distributions.categorical.Categorical
        orig_sample = Cat.sample
        rec = self

        def randint(*a, **k):
            out = orig_randint(*a, **k)
            rec.anchor_idx.append(out.clone().numpy().astype(np.int64))
            return out

        def sampler(samp_num, seg_num_list):
            out = orig_sampler(samp_num, seg_num_list)
            rec.neg_idx.append(np.asarray(out, dtype=np.int64))
            return out

        def sample(self_, sample_shape=torch.Size()):
            out = orig_sample(self_, sample_shape)
            rec.samp_class.append(out.clone().numpy().astype(np.int64))
            return out

        torch.randint = randint
        L.negative_index_sampler = sampler
        Cat.sample = sample
        try:
            yield self
        finally:
            torch.randint = orig_randint
            L.negative_index_sampler = orig_sampler
            Cat.sample = orig_sample


class StubNet(torch.nn.Module):
    """Stands in for DeepLabv3Plus_with_rep: returns pre-made (pred, rep) in call order."""

    def __init__(self, outputs):
        super().__init__()
        self.outputs = list(outputs)
        self.calls = 0

    def forward(self, x):
        out = self.outputs[self.calls % len(self.outputs)]
        self.calls += 1
        return out
