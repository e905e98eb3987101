#!/bin/bash
# quick loop: lift parity tests + bench line (+ optional ncu of the forward kernel when $1 = ncu)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lift_forward_kernel -s 1 -c 1 -f -o gpurun_out/prof_lift_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
tail -4 gpurun_out/pytest_gpu.log; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['value'], d['roofline']); print({k: d[k] for k in d if k.startswith('value_') or k in ('fwd_bwd','e2e')})"; tail -2 gpurun_out/bench.err
