#!/bin/bash
# re-entry check: parity tests, smoke, bench line, launch list, full capture of the column tile kernel and the layout pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lift_forward_cols_kernel -s 8 -c 1 -f -o gpurun_out/prof_lift_fwd_cols python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:finalize_clear -s 8 -c 1 -f -o gpurun_out/prof_finalize_clear python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
