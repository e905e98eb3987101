#!/bin/bash
# round-1 final evidence: tests, smoke, bench lines for all workloads, launch list, full captures of the kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
for w in cfg3_baseline cfg4_pon cfg2_static_lss; do timeout 600 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2>> gpurun_out/bench.err; done
timeout 600 python bench.py --steps 20 --warmup 5 --layout channels_last --no-cpu-baseline > gpurun_out/bench_channels_last.json 2>> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
W=cfg2_static_lss_b8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lift_forward_cols_kernel -s 2 -c 1 -f -o gpurun_out/prof_lift_fwd_cols python tools/ncu_target.py $W tile > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lift_forward_cols_kernel -s 2 -c 1 -f -o gpurun_out/prof_lift_fwd_cols_cfg3 python tools/ncu_target.py cfg3_baseline tile >> gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:finalize_tma -s 8 -c 1 -f -o gpurun_out/prof_finalize_tma python tools/ncu_target.py $W step >> gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lift_backward_kernel -s 1 -c 1 -f -o gpurun_out/prof_lift_bwd python tools/ncu_target.py $W bwd >> gpurun_out/ncu_full.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
