"""Timing legs that run the UNMODIFIED reference (WangChangqi98/CSS) through its own public API.

TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's `cpu_baseline` leg and `--impl reference` arm): nothing in css_b200/ imports
this, and the product path never runs through it.  The reference tree is looked up by oracle/ref_harness.py
(CSS_REFERENCE_ROOT, then baseline/_ref -- the git-ignored copy __graft_entry__.build() makes, which travels to the GPU box --
then /root/reference).

What is driven, exactly as SURVEY.md 8(c) / Appendix B describe and BASELINE.md section 3 asks:
  * stages 1 / 1' / 2 / 1b: `Model_mix.forward` / `Model_cross.forward` / `Model_ori_pseudo.forward`
    (generalframeworks/networks/ddp_model.py:8-239) with the DeepLab networks replaced by stubs that return pre-made
    (pred, rep) tensors and the PIL augmentation replaced by pass-throughs -- both are outside the path (north_star) -- so the
    call executes the inline blocks :104-118,:147-154 (:189-199,:230-237 / :36-37) verbatim;
  * stages 3 / 4 + backward: `Contrast_Loss(...)(rep, label, mask, prob, prototypes)` and `.backward()`
    (generalframeworks/loss/loss.py:66-149,410-418).
The forward of the model shells also up-samples the student logits twice (ddp_model.py:136,139: inputs of the supervised
losses, not of this path); that small extra cost is part of the reference's own call and is left in.
"""
import os
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_harness

_STRATEGY = {"mix": ("Model_mix", "batch_transform_2", "generate_cut_gather_2", 4),
             "cross": ("Model_cross", "batch_transform_3", "generate_cut_gather_3", 5),
             "ori": ("Model_ori_pseudo", "batch_transform", "generate_cut_gather", 3)}


def available():
    return ref_harness.reference_available()


class _Stub(torch.nn.Module):
    def __init__(self, outs):
        super().__init__()
        self.outs, self.i = outs, 0

    def forward(self, x):
        o = self.outs[self.i % len(self.outs)]
        self.i += 1
        return o


class ReferencePath:
    """The reference's classes for one workload, on `device` ('cpu' or 'cuda:0')."""

    def __init__(self, cfg, device="cpu"):
        L, M, _ = ref_harness.load_reference(device)
        import torchvision.models as tvm
        self.cfg, self.device, self.L, self.M = cfg, torch.device(device), L, M
        name, bt, cg, n = _STRATEGY[cfg["strategy"]]
        conf = {"Dataset": {"crop_size": (cfg["H"], cfg["W"]), "scale_size": (1.0, 1.0), "mix_mode": "none"}}
        kw = {} if cfg["strategy"] == "ori" else {"temp": cfg["temp"]}
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):            # the constructors print a banner
            self.model = getattr(M, name)(tvm.resnet50(), num_classes=cfg["C"], output_dim=256, config=conf, **kw)
        # module-level names, resolved at call time (ddp_model.py:6,121,127,132): pass-through augmentation
        setattr(M, bt, lambda *a, **k: a[:n])
        setattr(M, cg, lambda *a, **k: a[:n])
        self.model.to(self.device)

    def stage12(self, inp, b_sub):
        """One forward of the model shell on the first b_sub teacher images / 2*b_sub student images -> prob_all (or None)."""
        cfg, dev = self.cfg, self.device
        B = inp["rep_u"].shape[0]
        rep_all, pred_u, rep_u = inp["rep_all"], inp["pred_u"][:b_sub], inp["rep_u"][:b_sub]
        rep_l, rep_s = rep_all[:b_sub], rep_all[B:B + b_sub]
        self.model.ema_model = _Stub([(pred_u, rep_u)])
        self.model.model = _Stub([(pred_u, rep_l), (pred_u, rep_s)])
        img = torch.zeros(b_sub, 3, cfg["H"], cfg["W"], device=dev)
        with torch.no_grad():
            if cfg["strategy"] == "ori":
                out = self.model(img, img)
                return None
            out = self.model(img, img, inp["prototypes"])
        return out[-1]

    def loss_step(self, inp, prob, protos, q_sub):
        cfg = self.cfg
        crit = self.L.Contrast_Loss(num_queries=q_sub, num_negatives=cfg["Nn"], temp=cfg["temp"], strong_threshold=cfg["strong"], alpha=0.99)
        rep = inp["rep_all"].detach().clone().requires_grad_(True)
        loss = crit(rep, inp["label"], inp["mask"], prob, protos)
        loss.backward()
        return float(loss.item())


def _sync(dev):
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)


def _to(inp, dev):
    return {k: v.to(dev) for k, v in inp.items()}


def _full_prob(path, inp):
    """prob of the whole batch for the loss (an INPUT of the loss; computed once, outside every timed region)."""
    cfg = path.cfg
    if cfg["strategy"] == "ori":
        return inp["prob_ori"]
    with torch.no_grad():
        x = F.normalize(inp["rep_all"].permute(0, 2, 3, 1), dim=-1)
        p = F.normalize(inp["prototypes"], dim=-1)
        return F.softmax((x @ p.t()).permute(0, 3, 1, 2) / cfg["temp"], dim=1).contiguous()


def run_cpu(cfg, host_inputs, steps, warmup, budget_s):
    """`steps` timed steps of the reference on the host cores after `warmup`, inside about `budget_s` seconds: when a full step
    does not fit, a step covers b_sub of B teacher images in the model shell (scaled by B / b_sub: every statement there is
    linear in images) and q_sub of Q queries per class in the loss (the loss time is affine in Q: per-class fixed work F
    + per-query work, calibrated from two short runs; scaled as F + (t - F) * Q / q_sub)."""
    with ref_harness.cpu_cuda_shim():
        return _run_cpu(cfg, host_inputs, steps, warmup, budget_s)


def _run_cpu(cfg, host_inputs, steps, warmup, budget_s):
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    path = ReferencePath(cfg, "cpu")
    inp = {k: v.clone() for k, v in host_inputs.items()}
    B, Q = cfg["B"], cfg["Q"]
    N = 2 * B * cfg["h"] * cfg["w"]
    prob = _full_prob(path, inp)
    protos = inp["prototypes"].clone()
    np.random.seed(0)
    torch.manual_seed(0)

    def t_stage(b_sub):
        t0 = time.perf_counter()
        path.stage12(inp, b_sub)
        return time.perf_counter() - t0

    def t_loss(q_sub, p):
        t0 = time.perf_counter()
        path.loss_step(inp, prob, p, q_sub)
        return time.perf_counter() - t0

    t_stage(1)                                                    # page in, warm the thread pool
    a1 = t_stage(1)
    l4 = t_loss(4, protos.clone())
    l4 = t_loss(4, protos.clone())
    l16 = t_loss(16, protos.clone())
    c_q = max((l16 - l4) / 12.0, 1e-6)
    fixed = max(l4 - 4 * c_q, 0.0)
    total = max(steps + warmup, 1)
    per_step = max(budget_s / total, 0.05)
    b_sub, q_sub = 1, 4
    while b_sub < B and a1 * (2 * b_sub) + fixed + q_sub * c_q <= 0.5 * per_step:
        b_sub *= 2
    while q_sub < Q and a1 * b_sub + fixed + (2 * q_sub) * c_q <= per_step:
        q_sub *= 2
    b_sub, q_sub = min(b_sub, B), min(q_sub, Q)
    ts = []
    for i in range(total):
        a = t_stage(b_sub)
        l = t_loss(q_sub, protos)
        if i >= warmup:
            ts.append(a * (B / b_sub) + (l if q_sub == Q else fixed + max(l - fixed, 0.0) * (Q / q_sub)))
    t_step = sum(ts) / len(ts)
    full = b_sub == B and q_sub == Q
    sample = (f"UNMODIFIED reference on the host cores (torch {torch.__version__} CPU, {cores} threads): {cfg['strategy']} model shell "
              f"forward (stub networks, pass-through augmentation) on {b_sub} of {B} teacher + {2 * b_sub} of {2 * B} student images, "
              f"Contrast_Loss forward+backward on the full batch with {q_sub} of {Q} queries/class (Nn={cfg['Nn']}); "
              + ("nothing scaled: every step is the full workload" if full else
                 f"shell time scaled by {B}/{b_sub}, loss time t scaled as F + (t - F)*{Q}/{q_sub} with the calibrated per-step "
                 f"fixed cost F = {fixed:.2f} s (per-query cost {c_q * 1e3:.1f} ms)"))
    return dict(value=N / t_step, t_step=t_step, cores=cores, sample=sample, kind="reference", full=full)


def run_gpu_eager(cfg, host_inputs, steps=2, warmup=1, device="cuda:0"):
    """The same reference code, eager, on the B200 under the image's torch (context line: the stronger baseline, SURVEY.md
    2.2).  Full workload, nothing scaled; `steps` is small because one step takes seconds (>= 10^5 host synchronisations)."""
    dev = torch.device(device)
    path = ReferencePath(cfg, device)
    inp = _to(host_inputs, dev)
    N = 2 * cfg["B"] * cfg["h"] * cfg["w"]
    protos = inp["prototypes"].clone()
    np.random.seed(0)
    torch.manual_seed(0)
    ts, t_stage, t_loss = [], [], []
    for i in range(steps + warmup):
        _sync(dev)
        t0 = time.perf_counter()
        prob = path.stage12(inp, cfg["B"])
        if prob is None:
            prob = inp["prob_ori"]
        _sync(dev)
        t1 = time.perf_counter()
        path.loss_step(inp, prob, protos, cfg["Q"])
        _sync(dev)
        t2 = time.perf_counter()
        if i >= warmup:
            ts.append(t2 - t0)
            t_stage.append(t1 - t0)
            t_loss.append(t2 - t1)
    t_step = sum(ts) / len(ts)
    return dict(value=N / t_step, unit="pixels/s", ms_per_step=t_step * 1e3, ms_model_shell=sum(t_stage) / len(t_stage) * 1e3,
                ms_loss_fwd_bwd=sum(t_loss) / len(t_loss) * 1e3, steps=steps, warmup=warmup,
                sample=f"UNMODIFIED reference, eager torch {torch.__version__} on {torch.cuda.get_device_name(dev)}: full workload "
                       f"({cfg['strategy']} model shell with stub networks + Contrast_Loss forward+backward, Q={cfg['Q']}, Nn={cfg['Nn']}), "
                       f"wall clock with device synchronisation, {steps} timed step(s) after {warmup} warm-up")
