#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/ref_gpu_breakdown.py > gpurun_out/ref_gpu_breakdown.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg4_pon --no-cpu-baseline > gpurun_out/bench_cfg4.json 2>> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:finalize -s 8 -c 1 -f -o gpurun_out/prof_finalize python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/ref_gpu_breakdown.txt; cat gpurun_out/bench.json | cut -c1-400; tail -2 gpurun_out/bench.err
