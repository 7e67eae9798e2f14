#!/bin/bash
# Round-end evidence on one GPU: parity tests, the default bench line (both arms), bench lines of the other workloads, per-op
# timings, the configs[4] sweep, the launch list and ncu full-set captures.  gpurun --timeout 2400 -- 'bash tools/gpu_final.sh r02zz'
set -u
TAG=${1:-r02zz}; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench_voc321_mix.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err; echo "reference arm rc=$?"
for wl in voc321_ori city768_cross voc81_b1_loss voc321_mix_nhwc city768_cross_nhwc; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_${wl}.json 2>> $OUT/${TAG}_bench.err; echo "$wl rc=$?"
done
timeout 300 python tools/kbench.py --iters 20 > $OUT/${TAG}_kbench_voc321_fp32.txt 2>&1
timeout 300 python tools/kbench.py --iters 20 --workload city768_cross > $OUT/${TAG}_kbench_city768_fp32.txt 2>&1
timeout 300 python tools/kbench.py --iters 20 --bf16 > $OUT/${TAG}_kbench_voc321_bf16.txt 2>&1
timeout 300 python tools/kbench.py --iters 20 --workload voc321_mix_nhwc > $OUT/${TAG}_kbench_voc321_nhwc.txt 2>&1
timeout 600 python tools/sweep.py --out $OUT/${TAG}_sweep_1gpu.json > $OUT/${TAG}_sweep_1gpu.txt 2>&1; echo "sweep rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_launches.log 2>&1
for k in rep_pass_kernel score_ce grad_slab_kernel upsample_label_fuse_kernel class_sums_kernel rows_verify_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 3 -f -o $OUT/${TAG}_ncu_$k \
      python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_ncu_$k.log 2>&1
done
python - <<PY
import json
for wl in ("voc321_mix","voc321_ori","city768_cross","voc81_b1_loss","voc321_mix_nhwc","city768_cross_nhwc"):
    try:
        d=json.load(open("$OUT/${TAG}_bench_%s.json"%wl)); print(wl, round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,1), "M px/s")
    except Exception as e: print(wl, "ERR", e)
PY
