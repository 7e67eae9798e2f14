"""Reads .ncu-rep captures (here, no GPU needed) and writes the per-kernel summary the roofline numbers come from:
    python tools/ncu_summary.py gpurun_out/r02a_ncu_*.ncu-rep --out profiles/r02a_ncu_full_summary.csv [--traffic voc321_mix]
--traffic WORKLOAD also merges dram bytes per launch into profiles/kernel_traffic.json (what bench.py reports as `traffic`)."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
           "launch__grid_size", "launch__block_size", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        sys.stderr.write(out.stderr)
        return []
    rows = list(csv.reader(io.StringIO(out.stdout)))
    if len(rows) < 3:
        return []
    head, units = rows[0], rows[1]
    return [dict(zip(head, r)) for r in rows[2:]], dict(zip(head, units))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reps", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic", default=None)
    a = ap.parse_args()
    recs = []
    for rep in a.reps:
        got = raw_rows(rep)
        if not got:
            continue
        rows, units = got
        for r in rows:
            rec = {"capture": os.path.basename(rep), "kernel": r.get("Kernel Name", "")[:110], "id": r.get("ID", "")}
            for m in METRICS:
                if m in r:
                    rec[m + (" [" + units.get(m, "") + "]" if units.get(m) else "")] = r[m]
            recs.append(rec)
    keys = []
    for r in recs:
        for k in r:
            if k not in keys:
                keys.append(k)
    with open(a.out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for r in recs:
            w.writerow(r)
    print(f"{len(recs)} kernel instances -> {a.out}")
    if a.traffic:
        path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        wl = data.setdefault(a.traffic, {})

        def num(x):
            return float(str(x).replace(",", "")) if x not in (None, "") else 0.0

        def to_bytes(v, unit):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            return num(v) * scale
        agg = {}
        for r in recs:
            rd = [k for k in r if k.startswith("dram__bytes_read.sum")]
            wr = [k for k in r if k.startswith("dram__bytes_write.sum")]
            if not rd or not wr:
                continue
            unit_r = rd[0].split("[")[-1].rstrip("]") if "[" in rd[0] else "byte"
            unit_w = wr[0].split("[")[-1].rstrip("]") if "[" in wr[0] else "byte"
            name = r["kernel"]
            key = name.split("(")[0].split("<")[0].strip().split(" ")[-1]
            if key == "rep_pass_kernel":
                rows_arg = name.split("<")[1].split(">")[0].split(",")[1].strip() if "<" in name else "0"     # ROWS template argument
                key = "rep_pass_student" if rows_arg in ("true", "1") else "rep_pass_teacher"
            agg.setdefault(key, []).append(to_bytes(r[rd[0]], unit_r) + to_bytes(r[wr[0]], unit_w))
        for k, v in agg.items():
            wl[k] = int(sum(v) / len(v))
        json.dump(data, open(path, "w"), indent=1)
        print(f"traffic of {sorted(agg)} merged into {path}")


if __name__ == "__main__":
    main()
