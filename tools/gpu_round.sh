#!/bin/bash
# One GPU-box visit: parity tests, bench lines, per-op timings, launch list, ncu full-set captures of the path's kernels,
# and the dev microbenchmarks.  Everything lands in gpurun_out/<tag>_*; summaries are copied into profiles/ afterwards.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02a [tests|notests] [ncu|noncu]'
set -u
TAG=${1:-r02x}; DO_TESTS=${2:-tests}; DO_NCU=${3:-ncu}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/${TAG}_smi.csv 2>&1
if [ "$DO_TESTS" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_voc321_mix.json 2> $OUT/${TAG}_bench_voc321_mix.err; echo "bench rc=$?"
tail -c 1500 $OUT/${TAG}_bench_voc321_mix.json
timeout 300 python tools/kbench.py --iters 20 > $OUT/${TAG}_kbench_voc321_fp32.txt 2>&1
timeout 300 python tools/kbench.py --iters 20 --workload city768_cross > $OUT/${TAG}_kbench_city768_fp32.txt 2>&1
head -20 $OUT/${TAG}_kbench_voc321_fp32.txt
for b in dev_gather dev_nchw dev_fma; do
  if [ -x tools/dev/$b ]; then timeout 120 tools/dev/$b > $OUT/${TAG}_${b}.txt 2>&1; fi
done
if [ "$DO_NCU" = "ncu" ]; then
  # every launch of two eager steps with its device time (cold-cache, serialised: shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_launches.log 2>&1
  # full-set capture of the path's kernels on the shipped build (3 instances each after the warm-up launches)
  for k in rep_pass_kernel score_ce_kernel grad_slab_kernel upsample_label_fuse_kernel class_sums_kernel select_classify_kernel; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 3 -f -o $OUT/${TAG}_ncu_$k \
        python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_ncu_$k.log 2>&1
  done
  timeout 900 ncu --set full --clock-control none -k regex:score_ce_kernel -s 2 -c 2 -f -o $OUT/${TAG}_ncu_city_score_ce_kernel \
      python bench.py --workload city768_cross --steps 2 --warmup 1 --no-graph --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_ncu_city.log 2>&1
fi
ls -la $OUT | tail -40
