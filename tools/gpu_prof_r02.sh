#!/bin/bash
# round-2 profiling pass: full captures of the three hot kernels + launch list of the bench command
mkdir -p gpurun_out
W=${1:-cfg2_static_lss_b8}
T=${2:-r02}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lift_forward_cols_kernel -s 2 -c 1 -f -o gpurun_out/${T}_prof_fwd_cols python tools/ncu_target.py $W tile > gpurun_out/${T}_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lift_plan_kernel -s 2 -c 1 -f -o gpurun_out/${T}_prof_plan python tools/ncu_target.py $W tile >> gpurun_out/${T}_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lift_backward_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_bwd python tools/ncu_target.py $W bwd >> gpurun_out/${T}_ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu_bench.log 2>&1
ls -la gpurun_out | tail -8
