"""SASS of the shipped kernels (cuobjdump -sass of css_b200/libcss_b200.so), stripped to address + instruction, plus an opcode
histogram per kernel:  python tools/sass_listing.py --out profiles/r02u_sass  [--kernels rep_pass_kernel score_ce_kernel ...]"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = ["rep_pass_kernelILi6ELb0EfLb1E", "rep_pass_kernelILi6ELb1EfLb1E", "score_ce_kernelILb1ELb0ELb1ELb1EfE", "score_ce_bulk_kernelILb1ELb0E",
           "grad_slab_kernel", "upsample_label_fuse_kernelILb1ELi21E", "class_sums_kernelIfE", "rep_pass_nhwc_kernelILi3E", "rep_pass_tc_kernel"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "css_b200", "libcss_b200.so"))
    ap.add_argument("--out", required=True)
    ap.add_argument("--kernels", nargs="*", default=DEFAULT)
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    hist_lines = []
    with open(a.out + "_listing.txt", "w") as f:
        for fn in funcs:
            name = fn.split("\n", 1)[0].strip()
            if not any(k in name for k in a.kernels):
                continue
            demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            ops = collections.Counter()
            f.write(f"==== {demangled[:200]}\n")
            for line in fn.split("\n"):
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
                if not m:
                    continue
                ins = m.group(2).strip()
                f.write(f"{m.group(1)}  {ins}\n")
                op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
                ops[op.split(".")[0]] += 1
            top = ", ".join(f"{k} {v}" for k, v in ops.most_common(14))
            marks = {k: sum(v for o, v in ops.items() if re.match(k, o)) for k in ("FFMA2", "UTC.?MMA", "UTMALDG", "UBLKCP", "LDTM", "LDGSTS", "SYNCS", "MUFU")}
            hist_lines.append(f"{demangled[:110]}\n    {sum(ops.values())} instructions: {top}\n    markers: {marks}\n")
    with open(a.out + "_opcodes.txt", "w") as f:
        f.write("Opcode histograms of the shipped kernels (cuobjdump -sass css_b200/libcss_b200.so; listing in the _listing.txt twin).\n"
                "FFMA2 = packed fp32x2 FMA (sm_100), UTC?MMA (UTCHMMA ...) = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,\n"
                "SYNCS = mbarrier, LDGSTS = cp.async.\n\n" + "\n".join(hist_lines))
    print("".join(hist_lines))


if __name__ == "__main__":
    main()
