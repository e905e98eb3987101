#!/bin/bash
mkdir -p gpurun_out
for t in test_plan_gpu test_lift_gpu test_lift_backward_gpu; do timeout 600 python -m pytest tests/$t.py -q -x 2>&1 | tail -25 > gpurun_out/r02_c_$t.log; echo "== $t"; tail -3 gpurun_out/r02_c_$t.log; done
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02_c_bench.json 2> gpurun_out/r02_c_bench.err; tail -c 300 gpurun_out/r02_c_bench.err
export FIERY_B200_LIB=$PWD/fiery_b200/libfiery_b200_ab.so
for w in cfg2_static_lss_b8 cfg3_baseline; do
  echo "== $w"; timeout 600 python tools/ab_forward.py $w 0,2,6,4,1,8 2>&1 | grep -E "^tile|^chains|Error|error" | cut -c1-330
  timeout 300 python tools/ab_backward.py $w 2>&1 | grep -E "^bwd|Error|error" | cut -c1-300
done
