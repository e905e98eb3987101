#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
for c in 1 2 4; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-chunk $c 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e chunk $c', d['e2e'])" >> gpurun_out/e2e_chunks.log; done
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg3_baseline --no-cpu-baseline > gpurun_out/bench_cfg3.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg4_pon --no-cpu-baseline > gpurun_out/bench_cfg4.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg2_static_lss --no-cpu-baseline > gpurun_out/bench_cfg2_b1.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --layout channels_last --no-cpu-baseline > gpurun_out/bench_channels_last.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lift_forward_kernel -s 8 -c 1 -f -o gpurun_out/prof_lift_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:finalize -s 8 -c 1 -f -o gpurun_out/prof_finalize python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/e2e_chunks.log; tail -2 gpurun_out/bench.err
