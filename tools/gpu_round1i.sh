#!/bin/bash
# 2-GPU weak-scaling check of bench.py under torchrun (as the driver launches it), plus the reference arm
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $?" >> gpurun_out/bench_2gpu.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "exit $?" >> gpurun_out/bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?" >> gpurun_out/bench_ref.err
cat gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_1gpu.json | cut -c1-600; cat gpurun_out/bench_ref.json; tail -2 gpurun_out/bench_ref.err
