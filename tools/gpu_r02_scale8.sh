#!/bin/bash
N=${1:-8}
T=${2:-r02s8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
run --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; tail -c 200 gpurun_out/${T}_bench_${N}gpu.err
NCCL_DEBUG=INFO run --no-cpu-baseline --direction fwd_bwd --workload cfg3_baseline --head-dtype f16 > gpurun_out/${T}_bench_train_${N}gpu.json 2> gpurun_out/${T}_bench_train_${N}gpu.err; grep -m3 -E "NVLS|Connected all rings|Using network" gpurun_out/${T}_bench_train_${N}gpu.err | cut -c1-200
python - <<PY
import json
for f in ("gpurun_out/${T}_bench_${N}gpu.json", "gpurun_out/${T}_bench_train_${N}gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), d.get("details", {}).get("host"))
    except Exception as e:
        print(f, "ERR", e)
PY
