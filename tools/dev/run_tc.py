"""Dev harness: tcgen05 rep pass (css_sim_tc.cu) vs the FFMA2 rep pass (css_sim.cu) vs a float64 torch reference.
    python tools/dev/run_tc.py            (needs a B200)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import css_b200  # noqa: E402
from css_b200 import _lib, synth  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def ref64(rep, protos, temp, softmax):
    x = rep.double().permute(0, 2, 3, 1)
    x = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    p = protos.double()
    p = p / p.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    s = (x @ p.t()).permute(0, 3, 1, 2)
    return torch.softmax(s / temp, dim=1) if softmax else s


def main():
    lib = _lib.load()
    cases = [(8, 21, 81, 81), (16, 21, 81, 81), (4, 19, 193, 193), (8, 19, 193, 193), (1, 5, 7, 9), (3, 32, 20, 21), (2, 21, 33, 31)]
    for B, C, h, w in cases:
        d = synth.student_batch(B, C, h, w, seed=5 + C, block=4)
        rep = d["rep"].cuda()
        protos = (0.5 * d["centers"] + 0.3 * synth.warm_prototypes(C, seed=2, zero_rows=(C // 2,))).cuda()
        out = {}
        for name, flag in (("fma", 0), ("tc", 1)):
            lib.css_set_rep_pass_path(flag)
            sim = css_b200.ops.cos_sim_map(rep, protos)
            prob = css_b200.ops.proto_softmax_sim(rep, protos, 0.5)
            torch.cuda.synchronize()
            out[name] = (sim, prob, prob._css_rows.rows, prob._css_rows.norms)
        r_sim, r_prob = ref64(rep, protos, 0.5, False), ref64(rep, protos, 0.5, True)
        line = f"B={B} C={C} {h}x{w}: "
        for name in ("fma", "tc"):
            sim, prob, rows, norms = out[name]
            line += (f"[{name}] sim err {float((sim.double() - r_sim).abs().max()):.2e} prob err {float((prob.double() - r_prob).abs().max()):.2e} "
                     f"finite {bool(torch.isfinite(sim).all())} ")
        rows_eq = torch.equal(out["fma"][2], out["tc"][2])
        want_rows = rep.permute(0, 2, 3, 1).reshape(-1, 256)
        line += f"rows tc==fma {rows_eq} rows tc==rep {torch.equal(out['tc'][2], want_rows)} norms maxdiff {float((out['fma'][3] - out['tc'][3]).abs().max()):.2e} "
        line += f"tc-vs-fma sim {float((out['fma'][0] - out['tc'][0]).abs().max()):.2e} zero-row-sim {float(out['tc'][0][:, C // 2].abs().max()):.1e}"
        print(line, flush=True)
        if B * h * w >= 50000:
            for name, flag in (("fma", 0), ("tc", 1)):
                lib.css_set_rep_pass_path(flag)
                t_teacher = timeit(lambda: css_b200.ops.cos_sim_map(rep, protos))
                t_student = timeit(lambda: css_b200.ops.proto_softmax_sim(rep, protos, 0.5))
                print(f"    {name}: sim-only {t_teacher:7.1f} us   sim+softmax+rows {t_student:7.1f} us   ({rep.numel() * 4 / 1e6:.0f} MB map)", flush=True)
    lib.css_set_rep_pass_path(-1)


if __name__ == "__main__":
    main()
