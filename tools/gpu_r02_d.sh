#!/bin/bash
mkdir -p gpurun_out
for t in test_plan_gpu test_lift_gpu test_lift_backward_gpu test_bench_configs_gpu; do timeout 900 python -m pytest tests/$t.py -q -x 2>&1 | tail -25 > gpurun_out/r02_d_$t.log; echo "== $t"; tail -3 gpurun_out/r02_d_$t.log; done
timeout 300 python bench.py --no-extras > gpurun_out/r02_d_bench.json 2> gpurun_out/r02_d_bench.err; tail -c 300 gpurun_out/r02_d_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_d_bench_ref.json 2>> gpurun_out/r02_d_bench.err
