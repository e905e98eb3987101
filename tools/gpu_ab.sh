#!/bin/bash
# A/B of an experiment build (python -m fiery_b200.build with FIERY_NVCC_EXTRA=-DFIERY_COLS_AB): unit shapes of the column tile
# kernel, then number of frame-group chains x minimum tiles per group (AB_COMBOS = "chains:min_tiles ..."), forward (eager and
# graph) and backward, then the lift parity tests on the requested combinations (AB_TEST, same format)
mkdir -p gpurun_out
# the experiment build sits next to the in-tree library: FIERY_NVCC_EXTRA=-DFIERY_COLS_AB python -m fiery_b200.build --out fiery_b200/libfiery_b200_ab.so
export FIERY_B200_LIB=${FIERY_B200_LIB:-$PWD/fiery_b200/libfiery_b200_ab.so}
for w in cfg2_static_lss_b8 cfg3_baseline cfg4_pon; do
  echo "== $w"; timeout 600 python tools/ab_forward.py $w 2>&1 | grep -E "^tile|^chains|Error|error" | cut -c1-330
done
for combo in $AB_TEST; do
  IFS=: read ch mt <<< "$combo"
  echo "== parity tests with $ch chains, >= $mt tiles per group"
  FIERY_CHAINS=$ch FIERY_CHAIN_MIN_TILES=$mt timeout 600 python -m pytest tests/test_lift_gpu.py -m gpu -q --no-header -x 2>&1 | tail -3
done
