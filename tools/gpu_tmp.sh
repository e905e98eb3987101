#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lift_warp_gpu.py -m gpu -q --no-header 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_adjoint_nhwc_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_warp_adj python tools/ncu_target.py cfg3_baseline bwd_warped > gpurun_out/r02_ncu_wa.log 2>&1
tail -2 gpurun_out/r02_ncu_wa.log
