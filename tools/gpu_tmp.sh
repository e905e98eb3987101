N=8
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02k_bench_8gpu.json 2> gpurun_out/r02k_bench_8gpu.err; tail -c 300 gpurun_out/r02k_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r02k_bench_ref_8gpu.json 2>> gpurun_out/r02k_bench_8gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02k_bench_8gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "median", d.get("ms_per_step_median_rank0"), "max", d.get("ms_per_step_max_rank0"), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), "static", d["ms_per_step_static_rig"], "fb", d["fwd_bwd"]["ms_per_step"], d["details"]["host"])
r=json.loads(open("gpurun_out/r02k_bench_ref_8gpu.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["config"]==d["config"])
PY
