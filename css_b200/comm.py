"""The path's one exchange step over NVLink peer memory: sum of every rank's [C, D+1] class statistics on one node.

Each rank allocates a small communication buffer through the C ABI (`css_comm_alloc`), the CUDA IPC handles travel once
through `torch.distributed.all_gather_object`, every rank maps its peers' buffers, and from then on the sum is ONE kernel
launch per step (`css_stats_allreduce`: remote stores + flags + a rank-ordered sum, include/css_b200.h) instead of the
reference's two `concat_all_gather`s (loss.py:77,81; ddp_model.py:241-250) -- or an NCCL all-reduce, which stays the path for
process groups that span nodes or cannot share memory.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr, stream_ptr


class PeerStatsReducer:
    """Collective constructor: every rank of `group` must create it at the same point."""

    def __init__(self, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.local = None
        self.mapped = {}
        self.ok = False
        lib = _lib.load()
        if os.environ.get("CSS_B200_COMM_TIMEOUT_S"):
            check(lib.css_comm_set_timeout_ms(int(float(os.environ["CSS_B200_COMM_TIMEOUT_S"]) * 1000)), "css_comm_set_timeout_ms")
        handle = None
        with torch.cuda.device(self.device):
            buf = ctypes.c_void_p()
            if self.world <= 16 and lib.css_comm_alloc(self.world, ctypes.byref(buf)) == 0:
                self.local = buf.value
                raw = ctypes.create_string_buffer(64)
                if lib.css_comm_export(self.local, raw) == 0:
                    handle = raw.raw
        # (host name, pid, handle): peer memory only makes sense between processes of one node
        mine = (os.uname().nodename, os.getpid(), handle)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        good = all(h is not None and node == mine[0] for node, _, h in everyone)
        ptrs = []
        if good:
            with torch.cuda.device(self.device):
                for r, (_, pid, h) in enumerate(everyone):
                    if r == self.rank:
                        ptrs.append(self.local)
                        continue
                    if pid == os.getpid():                   # same process (threads as ranks): IPC cannot map its own allocation
                        good = False
                        break
                    out = ctypes.c_void_p()
                    if lib.css_comm_open(h, ctypes.byref(out)) != 0:
                        good = False
                        break
                    self.mapped[r] = out.value
                    ptrs.append(out.value)
        verdict = [None] * self.world
        dist.all_gather_object(verdict, bool(good), group=group)   # all ranks take the same path, or none does
        self.ok = all(verdict)
        if self.ok:
            self.peer_table = torch.tensor(ptrs, dtype=torch.int64).to(self.device)
        else:
            self.close()

    def timeouts(self):
        """Calls that gave up waiting for a peer so far (read from pinned host memory: no synchronisation)."""
        return _lib.load().css_comm_timeouts(self.local) if self.local is not None else 0

    def stats(self):
        """dict(calls, timeouts, wait_ns_total, push_ns_total, kernel_ns_total): blocking device read, for diagnostics outside timed regions."""
        out = (ctypes.c_ulonglong * 5)()
        with torch.cuda.device(self.device):
            check(_lib.load().css_comm_stats(self.local, out), "css_comm_stats")
        return dict(calls=int(out[0]), timeouts=int(out[1]), wait_ns_total=int(out[2]), push_ns_total=int(out[3]), kernel_ns_total=int(out[4]))

    def allreduce(self, class_stats, C, D):
        lib = _lib.load()
        n = lib.css_comm_timeouts(self.local)
        if n > 0:      # the step that timed out used this rank's local statistics; the ranks are out of step from here on
            raise RuntimeError(f"css_b200: the peer-memory exchange timed out {n} time(s): a rank did not reach "
                               "Contrast_Loss.forward in time (CSS_B200_COMM_TIMEOUT_S, default 600 s); the affected step "
                               "updated the prototypes from rank-local statistics only")
        with torch.cuda.device(self.device):
            check(lib.css_stats_allreduce(ptr(class_stats), self.local, ptr(self.peer_table), self.rank, self.world, C, D, stream_ptr()),
                  "css_stats_allreduce")
        return class_stats

    def close(self):
        lib = _lib.load()
        with torch.cuda.device(self.device):
            for p in self.mapped.values():
                lib.css_comm_close(p)
            self.mapped = {}
            if self.local is not None:
                torch.cuda.synchronize(self.device)
                lib.css_comm_free(self.local)
                self.local = None
        self.ok = False

    def __del__(self):
        try:
            if self.local is not None and torch.cuda.is_available():
                self.close()
        except Exception:
            pass
