#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lift_warp_gpu.py tests/test_warp.py tests/test_lift_gpu.py -m gpu -q --no-header -x 2>&1 | tail -25
