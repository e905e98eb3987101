#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lift_forward_kernel -s 1 -c 1 -f -o gpurun_out/prof_lift_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_forward_kernel -s 1 -c 1 -f -o gpurun_out/prof_warp_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
tail -12 gpurun_out/pytest_gpu.log; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['value'], d['roofline']); print(d.get('next_row_cumulative_warp'))"; tail -2 gpurun_out/bench.err
