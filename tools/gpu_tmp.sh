#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg3_baseline --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg3_warp.json 2> gpurun_out/r02_bench_cfg3_warp.err
tail -3 gpurun_out/r02_bench_cfg3_warp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_cfg3_warp.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('next_row_cumulative_warp'),indent=1))
print(json.dumps(d.get('next_row_depth_layer'),indent=1))
print(d['value'], d['ms_per_step'])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:finalize_warp_kernel -s 4 -c 1 -f -o gpurun_out/r02_prof_finalize_warp python tools/ncu_target.py cfg3_baseline step_warped > gpurun_out/r02_ncu_fw.log 2>&1
tail -2 gpurun_out/r02_ncu_fw.log
