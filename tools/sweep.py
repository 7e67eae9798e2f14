"""BASELINE.json configs[4]: contrastive-path sweep over rep resolution x num_queries x num_negatives on one GPU.
Device time of the whole path (CUDA graph replay, same step as bench.py) -> profiles/<name>.json

    python tools/sweep.py --out profiles/r01_sweep_1gpu.json [--full]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import css_b200  # noqa: E402
from bench import make_inputs, path_bytes  # noqa: E402


def time_path(cfg, iters=20):
    dev = torch.device("cuda")
    host = make_inputs(cfg, 0)
    t = {k: v.to(dev) for k, v in host.items()}
    protos = t["prototypes"].clone()
    crit = css_b200.Contrast_Loss(num_queries=cfg["Q"], num_negatives=cfg["Nn"], temp=cfg["temp"], strong_threshold=cfg["strong"],
                                  seed=1).to(dev)
    H, W = cfg["H"], cfg["W"]

    def step():
        css_b200.ops.pseudo_labels(t["rep_u"], t["pred_u"], protos, cfg["temp"], (H, W), fuse="mix")
        prob = css_b200.ops.proto_softmax_sim(t["rep_all"], protos, cfg["temp"])
        rep = t["rep_all"].detach().requires_grad_(True)
        loss = crit(rep, t["label"], t["mask"], prob, protos)
        (g,) = torch.autograd.grad(loss, rep)
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    N = 2 * cfg["B"] * cfg["h"] * cfg["w"]
    stream_b, gather_b = path_bytes(cfg, cfg["C"])
    return dict(ms_per_step=ms, pixels_per_s=N / (ms * 1e-3), bytes_alg=stream_b + gather_b,
                alg_gbs=(stream_b + gather_b) / (ms * 1e-3) / 1e9, loss=float(loss.item()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_1gpu.json"))
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    sizes = [65, 81, 97, 129, 161, 193]
    qs = [128, 256, 512, 1024] if a.full else [128, 256, 1024]
    nns = [256, 512, 1024, 2048] if a.full else [256, 512, 2048]
    res = []
    for h in sizes:
        for Q in qs:
            for Nn in nns:
                if not a.full and not (Q == 256 or Nn == 512):
                    continue                      # cross through the default point
                cfg = dict(B=8 if h <= 129 else 4, C=21, h=h, w=h, H=4 * h - 3, W=4 * h - 3, Q=Q, Nn=Nn, temp=0.5, strong=0.8,
                           weak=0.7, strategy="mix")
                r = time_path(cfg)
                r.update(rep=h, B=cfg["B"], Q=Q, Nn=Nn)
                res.append(r)
                print(f"rep {h:3d}^2 B={cfg['B']} Q={Q:4d} Nn={Nn:4d}: {r['ms_per_step']:.3f} ms  {r['pixels_per_s'] / 1e6:7.1f} Mpx/s  "
                      f"{r['alg_gbs']:8.0f} GB/s(alg)", flush=True)
                torch.cuda.empty_cache()
    json.dump(dict(note="whole path (teacher labels + fusion, student prob + rows, loss fwd, backward), mix strategy, C=21, fp32, 1 GPU, "
                        "device time per CUDA-graph replay", results=res), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
