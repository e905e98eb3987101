#!/bin/bash
mkdir -p gpurun_out
T=r02f
timeout 300 python -m pytest tests/test_bev_conv_gpu.py -q 2>&1 | tail -12 > gpurun_out/${T}_test_conv.log; tail -4 gpurun_out/${T}_test_conv.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
W=cfg2_static_lss_b8
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lift_backward_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_bwd python tools/ncu_target.py $W bwd > gpurun_out/${T}_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lift_forward_cols_kernel -s 2 -c 1 -f -o gpurun_out/${T}_prof_fwd python tools/ncu_target.py $W tile >> gpurun_out/${T}_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bev_conv7x7s2 -s 2 -c 1 -f -o gpurun_out/${T}_prof_conv python tools/ncu_target.py $W conv >> gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
