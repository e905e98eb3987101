#!/bin/bash
# A/B of an experiment build (python -m fiery_b200.build with FIERY_NVCC_EXTRA=-DFIERY_COLS_AB): unit shapes of the column tile
# kernel x layout passes x stream chains (AB_COMBOS = "tile:pass:ctas:chains ..."), then the lift parity tests on the
# requested combinations (AB_TEST, same format)
mkdir -p gpurun_out
timeout 600 python tools/ab_forward.py cfg2_static_lss_b8 2>&1 | grep -E "^tile|^pass|Error|error" | cut -c1-260
timeout 300 python tools/ab_forward.py cfg3_baseline 2>&1 | grep -E "^tile|^pass|Error|error" | cut -c1-260
timeout 300 python tools/ab_forward.py cfg4_pon 2>&1 | grep -E "^tile|^pass|Error|error" | cut -c1-260
timeout 300 python tools/ab_forward.py cfg2_static_lss 2>&1 | grep -E "^tile|^pass|Error|error" | cut -c1-260
for combo in $AB_TEST; do
  IFS=: read v f c ch <<< "$combo"
  echo "== parity tests with tile variant $v, layout pass $f, $c CTAs/SM, $ch chains"
  FIERY_COLS_VARIANT=$v FIERY_FINALIZE=$f FIERY_FINALIZE_CTAS=$c FIERY_CHAINS=$ch timeout 600 python -m pytest tests/test_lift_gpu.py -m gpu -q --no-header -x 2>&1 | tail -3
done
