mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lift_backward_gpu.py tests/test_plan_gpu.py -q -x -p no:logging 2>&1 | tail -5
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err || tail -5 gpurun_out/r02j_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02j_bench.json"))
print("step", round(d["ms_per_step"]*1e3,1), "fwd_bwd graph", round(d["fwd_bwd"]["ms_per_step"]*1e3,1), "bwd nchw", round(d["roofline_bwd"]["step_ms"]*1e3,1), "cl", round(d["roofline_bwd"]["channels_last_grad"]["step_ms"]*1e3,1))
PY
