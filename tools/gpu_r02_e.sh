#!/bin/bash
mkdir -p gpurun_out
for t in test_bev_conv_gpu test_train_gpu; do timeout 300 python -m pytest tests/$t.py -q -x 2>&1 | tail -30 > gpurun_out/r02_e_$t.log; echo "== $t"; tail -12 gpurun_out/r02_e_$t.log; done
timeout 300 python bench.py --direction fwd_bwd --workload cfg3_baseline --head-dtype f16 --no-cpu-baseline > gpurun_out/r02_e_bench_train.json 2> gpurun_out/r02_e_bench_train.err; tail -c 400 gpurun_out/r02_e_bench_train.err; cut -c1-600 gpurun_out/r02_e_bench_train.json
