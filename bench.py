#!/usr/bin/env python
"""bench.py -- contrastive-path pixels/sec of the CSS representation-space hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload voc321_mix|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the four-stage path over one synthetic batch per GPU (SURVEY.md 8(d) `t_path`):
    teacher  : cosine map of rep_u vs prototypes -> fused bilinear-up + softmax + max (rep & logit space) + mix fusion
    student  : prob_all = softmax(cos(rep_all, prototypes)/temp)
    loss fwd : selection/compaction, one streaming read of rep_all (class sums + pixel-major normalised copy),
               [all-reduce of per-class sums|counts when N > 1], prototype EMA, sampling + scoring + CE
    loss bwd : dense grad_rep (zero fill + anchor scatter)
The DeepLabv3+ backbone, data loading and augmentation are outside the path (BASELINE.json north_star).
`pixels` = rep-resolution pixels of rep_all per step = 2B*h*w per GPU.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]/[2] shape: VOC, 321x321 crops, B=8 per GPU, rep 81x81, 21 classes
    "voc321_mix": dict(B=8, C=21, h=81, w=81, H=321, W=321, Q=256, Nn=512, temp=0.5, strong=0.8, weak=0.7, strategy="mix"),
    "voc321_ori": dict(B=8, C=21, h=81, w=81, H=321, W=321, Q=256, Nn=512, temp=0.5, strong=0.97, weak=0.7, strategy="ori"),
    # BASELINE.json configs[3]: CityScapes, 768x768 crops, deep-stem rep 193x193, 19 classes, B=4 per GPU (YAML)
    "city768_cross": dict(B=4, C=19, h=193, w=193, H=769, W=769, Q=256, Nn=512, temp=0.5, strong=0.8, weak=0.7, strategy="cross"),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (loss only)
    "voc81_b1_loss": dict(B=1, C=21, h=81, w=81, H=321, W=321, Q=256, Nn=512, temp=0.5, strong=0.97, weak=0.7, strategy="ori"),
}
D = 256
METRIC = "contrastive_path_pixels_per_sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="voc321_mix", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-budget-s", type=float, default=float(os.environ.get("CSS_CPU_BUDGET_S", 150)))
    return ap.parse_args()


def make_inputs(cfg, rank):
    """Synthetic tensors of the workload (CPU); seed 3407 (+rank), SURVEY.md 8(d)."""
    import torch
    from css_b200 import synth
    B, C, h, w = cfg["B"], cfg["C"], cfg["h"], cfg["w"]
    seed = 3407 + 1000 * rank
    t = synth.teacher_batch(B, C, h, w, seed=seed)
    s = synth.student_batch(2 * B, C, h, w, seed=seed + 1, strategy=cfg["strategy"])
    protos = 0.5 * s["centers"] + 0.3 * synth.warm_prototypes(C, seed=3407, zero_rows=(C - 1,))
    return dict(rep_u=t["rep_u"], pred_u=t["pred_u"], rep_all=s["rep"], label=s["label"], mask=s["mask"],
                prob_ori=torch.softmax(s["logits"], dim=1), prototypes=protos)


def path_bytes(cfg, v_eff):
    """Algorithmic bytes per step per GPU; fp32 everywhere (s = 4)."""
    B, C, h, w, H, W, Q, Nn = (cfg[k] for k in ("B", "C", "h", "w", "H", "W", "Q", "Nn"))
    Nu, N, NU = B * h * w, 2 * B * h * w, B * H * W
    out_b = {"mix": 28, "cross": 24, "ori": 12}[cfg["strategy"]]
    if cfg["strategy"] == "ori":
        stream = Nu * 4 * C + NU * out_b + N * (2 * D * 4 + 2 * 4 * C + 4)
    else:
        stream = Nu * (D * 4 + 4 * C) + NU * out_b + N * (2 * D * 4 + 2 * 4 * C + 4)
    gather = v_eff * Q * (Nn + 1) * D * 4
    return stream, gather


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            hnd = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(hnd, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n[len("nvmlClocksEventReason"):] for n in dir(nv) if n.startswith("nvmlClocksEventReason")
                     and isinstance(getattr(nv, n), int)}
            if not names:
                names = {getattr(nv, n): n[len("nvmlClocksThrottleReason"):] for n in dir(nv)
                         if n.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, n), int)}
            self.ok = True
            while not self._stop_evt.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(hnd, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(hnd)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
                for bit, name in names.items():
                    if bit and (r & bit) and name not in ("None", "GpuIdle", "All"):
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report that no clocks were sampled
            self.err = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (numpy restatement of the reference's CPU path) on the host cores
# ------------------------------------------------------------------------------------------------------------------
_PROB_CACHE = {}


def _cached_prob(inp, protos, temp, n_sub):
    """prob of the images beyond the timed sub-batch (needed as an input of the loss, computed once, outside the timing;
    prototypes drift slowly under the EMA, which does not change the cost of anything that is timed)."""
    from oracle import css_oracle as O
    key = (id(inp["rep_all"]), n_sub)
    if key not in _PROB_CACHE:
        _PROB_CACHE[key] = O.proto_softmax_sim(inp["rep_all"][n_sub:], protos, temp)
    return _PROB_CACHE[key]


def cpu_path_step(cfg, inp, protos, b_sub, q_sub):
    """One pass of the path on CPU with the oracle port, on a bounded sample: stage 1/2 on `b_sub` of B teacher images,
    loss fwd+bwd with `q_sub` of Q queries per class.  Returns (t_stage12, t_loss)."""
    import numpy as np
    from oracle import css_oracle as O
    C, H, W, temp = cfg["C"], cfg["H"], cfg["W"], cfg["temp"]
    t0 = time.perf_counter()
    if cfg["strategy"] == "ori":
        O.cls_pseudo_label(inp["pred_u"][:b_sub], (H, W))
        prob = inp["prob_ori"]
    else:
        _, label_rep, _ = O.rep_pseudo_label(inp["rep_u"][:b_sub], protos, temp, (H, W))
        _, label_cls = O.cls_pseudo_label(inp["pred_u"][:b_sub], (H, W))
        if cfg["strategy"] == "mix":
            O.mix_fuse(label_cls, label_rep, C)
        t_half = time.perf_counter()
        n_sub = 2 * b_sub                                              # student prob_all is linear in pixels: time it on a
        prob_sub = O.proto_softmax_sim(inp["rep_all"][:n_sub], protos, temp)   # sub-batch and scale, compute the rest untimed
        t_prob = (time.perf_counter() - t_half) * (inp["rep_all"].shape[0] / n_sub)
        prob = prob_sub if n_sub == inp["rep_all"].shape[0] else np.concatenate(
            [prob_sub, _cached_prob(inp, protos, temp, n_sub)])
    t1 = time.perf_counter()
    O.contrast_loss(inp["rep_all"], inp["label"], inp["mask"], prob, protos, num_queries=q_sub, num_negatives=cfg["Nn"],
                    temp=temp, strong_threshold=cfg["strong"], alpha=0.99, want_grad=True)
    t2 = time.perf_counter()
    if cfg["strategy"] == "ori":
        return (t1 - t0), 0.0, (t2 - t1)
    return t_half - t0, t_prob, (t2 - t1)


def run_cpu(cfg, steps, warmup, budget_s, rank_inputs):
    """Times the oracle port for `steps` steps after `warmup`, inside `budget_s` seconds by shrinking the per-step
    sample (teacher images, queries per class) and scaling the time back to the full step."""
    import numpy as np
    import torch
    inp = {k: v.numpy() for k, v in rank_inputs.items()}
    protos = inp["prototypes"].copy()
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    B, Q = cfg["B"], cfg["Q"]
    N = 2 * B * cfg["h"] * cfg["w"]
    # calibrate: the loss time is affine in the queries per class, t(q) = F + q*c (per-class fixed work + per-query work), so two
    # small runs give F and c and a step timed with q_sub queries is scaled as F + (t - F) * Q / q_sub (not t * Q / q_sub,
    # which would bill the reference Q/q_sub times for its fixed work); stage 1/2 and the student prob are linear in images
    np.random.seed(0)
    torch.manual_seed(0)
    b_sub = 1
    cpu_path_step(cfg, inp, protos.copy(), b_sub, 2)                       # warm numpy / BLAS threads, page in the inputs
    a, p, l8 = cpu_path_step(cfg, inp, protos.copy(), b_sub, 8)
    _, _, l32 = cpu_path_step(cfg, inp, protos.copy(), b_sub, 32)
    c_q = max((l32 - l8) / 24.0, 1e-6)
    fixed = max(l8 - 8 * c_q, 0.0)
    total_steps = steps + warmup
    per_step_budget = max(budget_s / max(total_steps, 1), 0.05)
    q_sub = 8
    while b_sub < B and (a + p) * 2 + fixed + q_sub * c_q <= 0.6 * per_step_budget:
        a, p, b_sub = a * 2, p, b_sub * 2            # p is already reported scaled to the full batch; its cost grows with b_sub
    while q_sub < Q and a + fixed + (2 * q_sub) * c_q <= per_step_budget:
        q_sub *= 2
    b_sub, q_sub = min(b_sub, B), min(q_sub, Q)
    ts = []
    for i in range(total_steps):
        a, p, l = cpu_path_step(cfg, inp, protos, b_sub, q_sub)
        if i >= warmup:
            l_full = l if q_sub == Q else fixed + max(l - fixed, 0.0) * (Q / q_sub)
            ts.append(a * (B / b_sub) + p + l_full)       # p is already scaled to the full batch
    t_step = sum(ts) / max(len(ts), 1)
    sample = (f"oracle port (numpy); each step = stage 1/2 on {b_sub} of {B} teacher images + student prob on {2 * b_sub} of "
              f"{2 * B} images + loss fwd+bwd on the full batch with {q_sub} of {Q} queries/class (Nn={cfg['Nn']}); stage 1/2 and prob "
              f"times are scaled by B/{b_sub}; the loss time t is scaled as F + (t - F)*Q/{q_sub} with the calibrated per-step fixed "
              f"cost F = {fixed:.2f} s")
    return dict(value=N / t_step, t_step=t_step, cores=cores, sample=sample)


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    inp = make_inputs(cfg, 0)
    r = run_cpu(cfg, args.steps, args.warmup, args.cpu_budget_s, inp)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "pixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["t_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg),
        "cpu_baseline": {"value": r["value"], "unit": "pixels/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, cfg):
    return {"workload": f"{args.workload}: {cfg['strategy']} strategy, B={cfg['B']}/GPU, C={cfg['C']}, rep {cfg['h']}x{cfg['w']}x{D} "
                        f"-> crop {cfg['H']}x{cfg['W']}, Q={cfg['Q']}, Nn={cfg['Nn']}, temp={cfg['temp']}, strong={cfg['strong']} "
                        "(BASELINE configs[1] shape; four-stage path incl. rep-space label + fusion of configs[2])",
            "pixels_per_step_per_gpu": 2 * cfg["B"] * cfg["h"] * cfg["w"],
            "parallelism": f"batch-sharded x{args.gpus}, one sum of the [C,D+1] class sums|counts block per step",
            "l2": "no explicit flush: per-step working set (rep_u + rep_all + pixel-major copy + grad_rep ~ 400 MB) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        return reference_arm(args, cfg)

    import torch
    import torch.distributed as dist
    import css_b200
    from css_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B, C, h, w, H, W, Q, Nn = (cfg[k] for k in ("B", "C", "h", "w", "H", "W", "Q", "Nn"))
    temp, strategy = cfg["temp"], cfg["strategy"]
    N = 2 * B * h * w
    host = make_inputs(cfg, rank)
    gpu = {k: v.to(dev) for k, v in host.items()}
    protos = gpu["prototypes"].clone()
    crit = css_b200.Contrast_Loss(num_queries=Q, num_negatives=Nn, temp=temp, strong_threshold=cfg["strong"], alpha=0.99,
                                  seed=3407 + rank).to(dev)

    def step(t):
        """one pass of the path on device-resident tensors `t`; returns (loss, grad_rep)"""
        if strategy == "ori":
            css_b200.ops.cls_pseudo_label(t["pred_u"], (H, W))
            prob = t["prob_ori"]
        else:
            css_b200.ops.pseudo_labels(t["rep_u"], t["pred_u"], protos, temp, (H, W), fuse="mix" if strategy == "mix" else "none")
            prob = css_b200.ops.proto_softmax_sim(t["rep_all"], protos, temp)
        rep = t["rep_all"].detach().requires_grad_(True)
        loss = crit(rep, t["label"], t["mask"], prob, protos)
        (grad,) = torch.autograd.grad(loss, rep)
        return loss, grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        loss, grad = step(gpu)
    barrier()
    meta = crit.last["ws"].meta.cpu().numpy()
    V = int(meta[_lib.META_V])
    v_eff = sum(1 for k in range(V) if meta[_lib.META_N_HARD + meta[_lib.META_CLS_OF_SLOT + k]] > 0) if V > 1 else 0
    loss_value = float(loss.item())

    # ---- one step captured as a CUDA graph (the library is capturable: no host sync, caller-owned buffers) -------------
    def capture(tensors):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step(tensors)
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(g):
            out = step(tensors)
        return g, out

    use_graph = not args.no_graph
    graph = None
    if use_graph:
        try:
            graph, graph_out = capture(gpu)
            graph.replay()
            barrier()
        except Exception as e:   # e.g. a collective that cannot be captured on this stack: fall back to eager launches
            sys.stderr.write(f"bench: CUDA graph capture failed ({e!r}); timing eager launches\n")
            graph, use_graph = None, False

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            step(gpu)

    # ---- timed region: device-resident inputs ------------------------------------------------------------------------
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    crit.score_events = [] if graph is None else None
    launches_per_step = None
    if graph is not None:      # launches inside a replayed graph are not re-counted by the library: count one eager step
        l0 = lib.css_launch_count()
        step(gpu)
        launches_per_step = lib.css_launch_count() - l0
        barrier()
    launches0 = lib.css_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = lib.css_launch_count() - launches0 if graph is None else launches_per_step * args.steps
    clocks = sampler.stop()
    if graph is not None:      # live CUDA-event timing of the dominant kernel, same stream, eager launches of the same steps
        crit.score_events = []
        for _ in range(min(args.steps, 50)):
            step(gpu)
        barrier()
    score_ms = [a.elapsed_time(b) for a, b in crit.score_events]
    crit.score_events = None
    t_el = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_el, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_el.item()) / args.steps
    value = world * N / (ms_per_step * 1e-3)

    # ---- e2e: host buffers, H2D of every input + D2H of the loss inside the timed region -----------------------------------
    keys = ["rep_u", "pred_u", "rep_all", "label", "mask"] if strategy != "ori" else ["pred_u", "rep_all", "label", "mask", "prob_ori"]
    pinned = {k: host[k].pin_memory() for k in keys}
    # two staging sets: the H2D copy of step i+1 (copy stream) overlaps the kernels of step i; every step still uploads all
    # of its inputs from pinned host memory and reads its loss back
    staged = [{k: torch.empty_like(gpu[k]) for k in keys} for _ in range(2)]
    for st_set in staged:
        for k in gpu:
            if k not in keys:
                st_set[k] = gpu[k]
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]      # inputs of the set have landed
    done = [torch.cuda.Event(), torch.cuda.Event()]       # the step that read the set has finished

    e2e_graphs = [None, None]
    if use_graph:
        try:
            e2e_graphs = [capture(staged[0]), capture(staged[1])]
        except Exception as e:
            sys.stderr.write(f"bench: e2e graph capture failed ({e!r}); eager\n")
            e2e_graphs = [None, None]

    def upload(i):
        s_ = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[s_])
            for k in keys:
                staged[s_][k].copy_(pinned[k], non_blocking=True)
            ready[s_].record(copy_stream)

    def e2e_run(n):
        for ev in done:
            ev.record()
        upload(0)
        for i in range(n):
            s_ = i % 2
            if i + 1 < n:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[s_])
            if e2e_graphs[s_] is not None:
                e2e_graphs[s_][0].replay()
                loss = e2e_graphs[s_][1][0]
            else:
                loss, grad = step(staged[s_])
            done[s_].record()
            float(loss.item())                # D2H read of the step's result (synchronises)

    e2e_run(4)
    barrier()
    e0.record()
    e2e_run(args.e2e_steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t_e2e = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * N / (float(t_e2e.item()) / args.e2e_steps * 1e-3)

    # ---- roofline of the dominant kernel (css_score_ce: per-query row gather + CE + d/d anchor) --------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    stream_b, gather_b = path_bytes(cfg, v_eff)
    score_avg_ms = sum(score_ms) / max(len(score_ms), 1)
    achieved = gather_b / (score_avg_ms * 1e-3) / 1e9 if score_avg_ms > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "score_ce_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "score_ce_kernel (css_score_ce)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "dram_gbs": (traffic / (score_avg_ms * 1e-3) / 1e9) if (traffic and score_avg_ms > 0) else None,
                "l2_served_frac": (1.0 - traffic / gather_b) if (traffic and gather_b) else None,
                "algorithmic_bytes_per_launch": gather_b, "avg_launch_ms": score_avg_ms, "share_of_step": score_avg_ms / ms_per_step,
                "note": "achieved counts LOGICAL 1 KB row gathers (SURVEY.md 8(d)); the 107 MB pixel-major copy is mostly L2 "
                        "resident at this size, so frac > 1 is L2 service, not HBM: `traffic` (ncu dram bytes per launch) and "
                        "`dram_gbs` say what really reached HBM; the kernel is bound by the L2->SM gather path (~17-19 TB/s "
                        "measured ceiling for random 1 KB rows, tools/dev/dev_gather.cu)",
                "path": {"bytes_stream": stream_b, "bytes_gather": gather_b,
                         "achieved_gbs": (stream_b + gather_b) / (ms_per_step * 1e-3) / 1e9,
                         "frac": (stream_b + gather_b) / (ms_per_step * 1e-3) / 1e9 / peak,
                         "compulsory_gbs": stream_b / (ms_per_step * 1e-3) / 1e9,
                         "compulsory_frac": stream_b / (ms_per_step * 1e-3) / 1e9 / peak}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # one bounded sample of the same workload on the host cores (oracle port), ~10-30 s
        r = run_cpu(cfg, 1, 0, 25.0, host)
        cpu = {"value": r["value"], "unit": "pixels/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pixels/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, cfg), "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "pixels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": float(t_e2e.item()) / args.e2e_steps, "steps": args.e2e_steps},
            "gpu_launches": int(launches) * world, "launch_mode": "cuda_graph" if graph is not None else "eager",
            "exchange": {"peer": "css_stats_allreduce over NVLink peer memory (one launch, rank-ordered sum)",
                         "nccl": "torch.distributed all_reduce", "none": "single process"}[crit.exchange_mode()],
            "clocks": clocks, "loss": loss_value, "present_classes": V, "scored_classes": v_eff,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # graphs that captured NCCL work must be released before the communicator goes away; then leave without the
        # interpreter's atexit teardown, which can deadlock on a communicator that was used inside a captured graph
        graph = None
        e2e_graphs = [None, None]
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
