#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py "tests/test_lift_gpu.py::test_lift_is_a_dispatcher_operator" -q -x 2>&1 | tail -40 > gpurun_out/r02t_test.log; tail -40 gpurun_out/r02t_test.log | cut -c1-220
