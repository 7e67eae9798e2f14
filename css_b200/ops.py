"""Functional ops of the representation-space hot path: thin wrappers that validate tensors, allocate outputs with
torch's caching allocator and call the C ABI (include/css_b200.h) on the current CUDA stream.

Every op raises on non-CUDA tensors: there is no CPU fallback (the CPU restatement lives in oracle/ and is test
infrastructure only).  Reference lines are relative to the reference root (WangChangqi98/CSS).
"""
import os

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _cuda_f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"css_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"css_b200: `{name}` must be float32, got {t.dtype}")
    return t.detach().contiguous()


def is_channels_last(t):
    """True for a float32 [B,256,h,w] map whose memory is [pixel][channel] (torch.channels_last) and 16-byte aligned: the layout
    the channels-last kernels (css_rep_pass_nhwc, css_grad_scatter_nhwc) take without any copy."""
    return (t.dim() == 4 and t.dtype == torch.float32 and t.shape[1] == _lib.D and not t.is_contiguous()
            and t.is_contiguous(memory_format=torch.channels_last) and t.data_ptr() % 16 == 0)


def rows_view(rep):
    """[N, 256] view of a channels-last map (no copy): row p = pixel id p."""
    B, D, h, w = rep.shape
    return rep.detach().permute(0, 2, 3, 1).reshape(B * h * w, D)


def _cuda_rep(t, name="rep"):
    """Representation maps may be float32 or bfloat16 (BASELINE north_star); everything downstream is fp32."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"css_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype == torch.float32:
        return t.detach().contiguous(), _lib.DTYPE_F32
    if t.dtype == torch.bfloat16:
        return t.detach().contiguous(), _lib.DTYPE_BF16
    raise RuntimeError(f"css_b200: `{name}` must be float32 or bfloat16, got {t.dtype}")


def _proto_scratch(device):
    return torch.empty(2 * _lib.D * _lib.CMAX, device=device, dtype=torch.float32)


def cos_sim_map(rep, prototypes):
    """Cosine similarity of every pixel against every class prototype: [B,D,h,w] x [C,D] -> [B,C,h,w].
    Replaces normalize + mm + reshape/permute of ddp_model.py:104-110."""
    return _sim(rep, prototypes, _lib.SIM_COS, 1.0)


def proto_softmax_sim(rep_all, prototypes, temp, with_rows=True):
    """prob_all = softmax(cos(rep_all, prototypes) / temp) at rep resolution (the hard-anchor indicator).
    Replaces ddp_model.py:147-154 (Model_mix) / :230-237 (Model_cross).

    The same single read of rep_all also writes the pixel-major copy + norms the contrastive loss needs; they ride on
    the returned tensor (`prob._css_rows`) so that Contrast_Loss.forward(rep_all, ..., prob, ...) does not read rep_all
    a second time.  Purely an optimisation: the loss re-derives them whenever the attachment does not match its `rep`."""
    return _sim(rep_all, prototypes, _lib.SIM_SOFTMAX, float(temp), with_rows)


class RowsCache:
    """Pixel-major copy of one representation map, riding on the `prob` tensor it was produced with.

    match(rep) decides how the loss may use it:
      "same"   -- `rep` is the very tensor the rows were read from (same storage address, same version counter).  The cache
                  keeps a detached alias of that tensor alive, so the caching allocator cannot hand its address to another
                  tensor while the cache exists;
      "verify" -- `rep` is another tensor of the same shape that was never modified in place (version 0), e.g. the clone
                  DistributedDataParallel(find_unused_parameters=True) makes of every output that requires grad
                  (mix_label.py:76-77).  The rows are then checked against `rep` ON THE DEVICE (css_rows_refresh: sampled
                  bit-for-bit comparison, and a rows-only pass that runs only if it failed) -- no host synchronisation;
      "miss"   -- anything else: the loss reads `rep` again.
    CSS_B200_ROWS_CACHE=strict disables "verify" (identity only)."""

    def __init__(self, rep, rows, norms):
        self.src = rep.detach()
        self.version = rep._version
        self.rows, self.norms = rows, norms

    def match(self, rep):
        src = self.src
        if rep.shape != src.shape or rep.dtype != src.dtype or rep.device != src.device or rep.stride() != src.stride():
            return "miss"
        if rep.data_ptr() == src.data_ptr() and rep._version == self.version and src._version == self.version:
            return "same"
        if rep._version == 0 and os.environ.get("CSS_B200_ROWS_CACHE", "verify") != "strict":
            return "verify"
        return "miss"

    def rebind(self, rep):
        """After css_rows_refresh the rows are those of `rep` (verified equal, or rewritten from it)."""
        self.src = rep.detach()
        self.version = rep._version


def rep_norms_nhwc(rep):
    """||x_p|| of a channels-last map (its rows are the map itself): one streaming read."""
    B, D, h, w = rep.shape
    norms = torch.empty(B * h * w, device=rep.device, dtype=torch.float32)
    scratch = _proto_scratch(rep.device)
    lib = _lib.load()
    with torch.cuda.device(rep.device):
        _timed_rep_pass("norms_nhwc",
                        lambda: check(lib.css_rep_pass_nhwc(ptr(rep), _lib.DTYPE_F32, None, ptr(scratch), B, 1, D, h, w, _lib.SIM_COS, 1.0, None,
                                                            ptr(norms), stream_ptr()), "css_rep_pass_nhwc"))
    return norms


def rep_rows(rep):
    """Pixel-major copy rows [N,256] (same dtype as the map) + norms [N] f32 of an NCHW representation map (one streaming read)."""
    rep, dt = _cuda_rep(rep)
    B, D, h, w = rep.shape
    rows = torch.empty((B * h * w, D), device=rep.device, dtype=rep.dtype)      # rows keep the map's dtype (bf16 is lossless)
    norms = torch.empty(B * h * w, device=rep.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(rep.device):
        _timed_rep_pass("rows_only",
                        lambda: check(lib.css_rep_pass(ptr(rep), dt, None, None, B, 1, D, h, w, _lib.SIM_COS, 1.0, None, ptr(rows), ptr(norms),
                                                       stream_ptr()), "css_rep_pass"))
    return rows, norms


# set to a list to collect (start, end, label) CUDA events around every css_rep_pass launch (bench.py's live kernel timing)
rep_pass_events = None


def _timed_rep_pass(label, launch):
    ev = rep_pass_events
    if ev is None:
        return launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    ev.append((e0, e1, label))


def _sim_nhwc(rep, prototypes, mode, temp, with_rows):
    """Channels-last map: similarities straight from the map (TMA + tcgen05, css_rep_pass_nhwc); the map is its own row table."""
    rep = rep.detach()
    prototypes = _cuda_f32(prototypes, "prototypes")
    B, D, h, w = rep.shape
    C = prototypes.shape[0]
    if prototypes.shape[1] != D:
        raise RuntimeError("css_b200: prototypes must be [C, D]")
    out = torch.empty((B, C, h, w), device=rep.device, dtype=torch.float32)
    norms = torch.empty(B * h * w, device=rep.device, dtype=torch.float32) if with_rows else None
    scratch = _proto_scratch(rep.device)
    lib = _lib.load()
    with torch.cuda.device(rep.device):
        _timed_rep_pass("student_nhwc" if with_rows else "teacher_nhwc",
                        lambda: check(lib.css_rep_pass_nhwc(ptr(rep), _lib.DTYPE_F32, ptr(prototypes), ptr(scratch), B, C, D, h, w, mode, temp,
                                                            ptr(out), ptr(norms), stream_ptr()), "css_rep_pass_nhwc"))
    if with_rows:
        out._css_rows = RowsCache(rep, rows_view(rep), norms)
    return out


def _sim(rep, prototypes, mode, temp, with_rows=False):
    if isinstance(rep, torch.Tensor) and rep.is_cuda and is_channels_last(rep):
        return _sim_nhwc(rep, prototypes, mode, temp, with_rows)
    rep, dt = _cuda_rep(rep)
    prototypes = _cuda_f32(prototypes, "prototypes")
    B, D, h, w = rep.shape
    C = prototypes.shape[0]
    if prototypes.shape[1] != D:
        raise RuntimeError("css_b200: prototypes must be [C, D]")
    out = torch.empty((B, C, h, w), device=rep.device, dtype=torch.float32)
    scratch = _proto_scratch(rep.device)
    rows = norms = None
    if with_rows:
        rows = torch.empty((B * h * w, D), device=rep.device, dtype=rep.dtype)  # rows keep the map's dtype (bf16 is lossless)
        norms = torch.empty(B * h * w, device=rep.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(rep.device):
        _timed_rep_pass("student" if with_rows else "teacher",
                        lambda: check(lib.css_rep_pass(ptr(rep), dt, ptr(prototypes), ptr(scratch), B, C, D, h, w, mode, temp, ptr(out),
                                                       ptr(rows), ptr(norms), stream_ptr()), "css_rep_pass"))
    if with_rows:
        out._css_rows = RowsCache(rep, rows, norms)
    return out


def upsample_label_fuse(sim, logits, temp, out_hw, fuse="none"):
    """Fused bilinear(align_corners=True) up-sampling + softmax + max of `sim` ([B,C,h,w] cosine map, scaled by 1/temp)
    and/or `logits` ([B,C,h,w]) at crop resolution, plus the mix-label fusion.  Replaces ddp_model.py:111-118.
    Returns a dict with conf_rep/label_rep (if sim), conf_cls/label_cls (if logits), fused (if fuse == 'mix')."""
    src = sim if sim is not None else logits
    if src is None:
        raise RuntimeError("css_b200: upsample_label_fuse needs sim and/or logits")
    sim = _cuda_f32(sim, "sim") if sim is not None else None
    logits = _cuda_f32(logits, "logits") if logits is not None else None
    B, C, h, w = src.shape
    H, W = int(out_hw[0]), int(out_hw[1])
    dev = src.device
    out = {}
    if sim is not None:
        out["conf_rep"] = torch.empty((B, H, W), device=dev, dtype=torch.float32)
        out["label_rep"] = torch.empty((B, H, W), device=dev, dtype=torch.int64)
    if logits is not None:
        if logits.shape != src.shape:
            raise RuntimeError("css_b200: sim and logits must have the same shape")
        out["conf_cls"] = torch.empty((B, H, W), device=dev, dtype=torch.float32)
        out["label_cls"] = torch.empty((B, H, W), device=dev, dtype=torch.int64)
    mode = _lib.FUSE_NONE
    if fuse == "mix":
        mode = _lib.FUSE_MIX
        out["fused"] = torch.empty((B, H, W), device=dev, dtype=torch.float32)
    elif fuse != "none":
        raise RuntimeError("css_b200: fuse must be 'none' or 'mix'")
    lib = _lib.load()
    with torch.cuda.device(dev):
        check(lib.css_upsample_label_fuse(ptr(sim), ptr(logits), float(temp), mode, B, C, h, w, H, W,
                                          ptr(out.get("conf_rep")), ptr(out.get("label_rep")), ptr(out.get("conf_cls")),
                                          ptr(out.get("label_cls")), ptr(out.get("fused")), stream_ptr()),
              "css_upsample_label_fuse")
    return out


def rep_pseudo_label(rep_u, prototypes, temp, out_hw):
    """(conf_rep f32[B,H,W], label_rep i64[B,H,W]): representation-space pseudo label. ddp_model.py:104-112."""
    o = upsample_label_fuse(cos_sim_map(rep_u, prototypes), None, temp, out_hw)
    return o["conf_rep"], o["label_rep"]


def cls_pseudo_label(pred_u, out_hw):
    """(conf_cls f32[B,H,W], label_cls i64[B,H,W]): logit-space pseudo label. ddp_model.py:113-114 (:36-37, :198-199)."""
    o = upsample_label_fuse(None, pred_u, 1.0, out_hw)
    return o["conf_cls"], o["label_cls"]


def pseudo_labels(rep_u, pred_u, prototypes, temp, out_hw, fuse="none"):
    """The whole teacher-side block in two launches: similarity map, then one fused up-sample/label/fuse kernel.
    Replaces ddp_model.py:104-118 (fuse='mix') and :189-199 (fuse='none')."""
    return upsample_label_fuse(cos_sim_map(rep_u, prototypes), pred_u, temp, out_hw, fuse)


def mix_fuse(label_cls, label_rep, num_classes):
    """Stand-alone mix fusion of two given label maps (ddp_model.py:115-118); the fused kernel above is the hot path."""
    if not label_cls.is_cuda:
        raise RuntimeError("css_b200: `label_cls` must be a CUDA tensor (no CPU fallback)")
    return torch.where(label_cls == label_rep, label_cls, torch.full_like(label_cls, 255)).float()


def threshold_glue(train_l_label, train_u_aug_label, train_u_aug_logits_cls, weak_threshold, num_class, out_hw, strategy):
    """Fused weak-threshold mask + one-hot + nearest down-sampling.
    Replaces mix_label.py:175-183 ('mix'), cross_label.py:178-185 ('cross'), ori_pseudo.py:171-178 ('ori').
    Returns (label_all [2B,C,h,w] f32, mask_all [2B,1,h,w] f32)."""
    for t, n in ((train_l_label, "train_l_label"), (train_u_aug_label, "train_u_aug_label"),
                 (train_u_aug_logits_cls, "train_u_aug_logits_cls")):
        if not t.is_cuda:
            raise RuntimeError(f"css_b200: `{n}` must be a CUDA tensor (no CPU fallback)")
    ll = train_l_label.long().contiguous()
    lu = train_u_aug_label.long().contiguous()
    cu = train_u_aug_logits_cls.float().contiguous()
    B, H, W = ll.shape
    h, w = int(out_hw[0]), int(out_hw[1])
    label_all = torch.empty((2 * B, num_class, h, w), device=ll.device, dtype=torch.float32)
    mask_all = torch.empty((2 * B, 1, h, w), device=ll.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(ll.device):
        check(lib.css_threshold_glue(ptr(ll), ptr(lu), ptr(cu), float(weak_threshold), 1 if strategy == "mix" else 0,
                                     B, num_class, H, W, h, w, ptr(label_all), ptr(mask_all), stream_ptr()),
              "css_threshold_glue")
    return label_all, mask_all
