#!/bin/bash
mkdir -p gpurun_out
T=r02g
timeout 1500 python -m pytest tests -m gpu -q -p no:logging --tb=short 2>&1 | grep -v "^DEBUG" | tail -40 > gpurun_out/${T}_pytest.log; tail -25 gpurun_out/${T}_pytest.log | cut -c1-250
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
