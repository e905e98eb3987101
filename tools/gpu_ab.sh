#!/bin/bash
# A/B of the layout pass: separate re-zeroing kernels vs the fused 32-byte variant
mkdir -p gpurun_out
for v in separate fused; do
  FIERY_FINALIZE=$v timeout 600 python -m pytest tests/test_lift_gpu.py -m gpu -q --no-header -x 2>&1 | tail -2
  FIERY_FINALIZE=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$v.json')); print('$v', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['lift_plus_finalize_ms'], d['value_eager'])" || tail -5 gpurun_out/bench_$v.err
done
FIERY_FINALIZE=fused timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "finalize|clear|lift_forward" gpurun_out/launches_fused.csv | awk -F, '{print $5, $NF}' | tail -8
