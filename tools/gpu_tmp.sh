#!/bin/bash
mkdir -p gpurun_out
export FIERY_B200_LIB=$PWD/fiery_b200/libfiery_b200_ab.so
for f in 0 1 2 4 3 6 7; do echo "skip=$f: $(FIERY_DL_SKIP=$f PYTHONPATH=. timeout 200 python tools/bench_depth_layer.py 8 fp16 2>&1 | tail -1)"; done
for fr in 4 16 32; do echo "frames=$fr: $(FIERY_DL_SKIP=0 PYTHONPATH=. timeout 200 python tools/bench_depth_layer.py $fr fp16 2>&1 | tail -1)"; done
