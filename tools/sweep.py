"""BASELINE.json configs[4]: contrastive-path sweep over rep resolution x num_queries x num_negatives on 1 or N GPUs.
Device time of the whole path (CUDA graph replay, same step as bench.py) -> gpurun_out/<name>.json

    python tools/sweep.py --out gpurun_out/sweep_1gpu.json [--full]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py --out gpurun_out/sweep_8gpu.json
Under torchrun every rank runs the same configurations on its own shard (weak scaling: per-GPU batch fixed); the step includes
the class-statistics exchange, times are the max over ranks and throughput is the sum over ranks.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import css_b200  # noqa: E402
from bench import make_inputs, path_bytes  # noqa: E402


RANK = int(os.environ.get("RANK", 0))
WORLD = int(os.environ.get("WORLD_SIZE", 1))


def time_path(cfg, iters=20):
    dev = torch.device("cuda", torch.cuda.current_device())
    host = make_inputs(cfg, RANK)
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
    if WORLD > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if WORLD > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    N = WORLD * 2 * cfg["B"] * cfg["h"] * cfg["w"]
    stream_b, gather_b = path_bytes(cfg, cfg["C"])
    out = dict(ms_per_step=ms, pixels_per_s=N / (ms * 1e-3), bytes_alg=stream_b + gather_b,
               alg_gbs=(stream_b + gather_b) / (ms * 1e-3) / 1e9, loss=float(loss.item()), exchange=crit.exchange_mode())
    del g                      # a graph that captured the exchange must go before its buffers do
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_1gpu.json"))
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    if WORLD > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
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
                if RANK == 0:
                    print(f"rep {h:3d}^2 B={cfg['B']} Q={Q:4d} Nn={Nn:4d}: {r['ms_per_step']:.3f} ms  {r['pixels_per_s'] / 1e6:7.1f} Mpx/s  "
                          f"{r['alg_gbs']:8.0f} GB/s(alg per GPU)", flush=True)
                torch.cuda.empty_cache()
    if RANK == 0:
        json.dump(dict(note=f"whole path (teacher labels + fusion, student prob + rows, loss fwd, backward), mix strategy, C=21, fp32, "
                            f"{WORLD} GPU(s), weak scaling, device time per CUDA-graph replay (max over ranks), pixels/s summed over ranks",
                       results=res), open(a.out, "w"), indent=1)
    if WORLD > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
