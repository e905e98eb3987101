#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_depth_layer_gpu.py -m gpu -q --no-header 2>&1 | tail -25
timeout 200 env PYTHONPATH=. python tools/bench_depth_layer.py 2>&1 | tail -5
