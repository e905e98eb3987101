#!/bin/bash
# parity tests, then the bench line (+ ncu of the forward kernel when $1 = ncu)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-warp}; do
  FIERY_LIFT_FORWARD=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$v.json')); print('$v', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['value_eager'])" || tail -5 gpurun_out/bench_$v.err
done
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lift_forward_cols -s 1 -c 1 -f -o gpurun_out/prof_lift_fwd_cols python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
