#!/bin/bash
# VoxelsSumming drop-in: parity tests + timing of apply() on one frame's rows next to an unrelated big allocation pattern
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_voxels_summing_gpu.py -m gpu -q --no-header 2>&1 | tail -2
timeout 300 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
from fiery_b200.geometry import VoxelsSumming
from fiery_b200.lift import LiftSplat
from fiery_b200.synthetic import CONFIGS, make_calibration
cfg = CONFIGS["cfg2_static_lss"]; dev = torch.device("cuda:0")
K, E = make_calibration(cfg, seed=100)
lift = LiftSplat.from_config(cfg).to(dev)
idx, valid, pillar = lift.point_indices(torch.from_numpy(K).to(dev), torch.from_numpy(E).to(dev))
keep = valid[0]; ranks = pillar[0][keep].long(); order = ranks.argsort(); ranks = ranks[order]; geo = idx[0][keep][order]
feats = torch.randn(ranks.numel(), 64, device=dev)
big = [torch.empty(256 << 20, dtype=torch.uint8, device=dev) for _ in range(8)]   # memory pressure like the bench
flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
for _ in range(3): VoxelsSumming.apply(feats, geo, ranks)
ts = []
for _ in range(20):
    flush.fill_(1.0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); VoxelsSumming.apply(feats, geo, ranks); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
print("VoxelsSumming.apply rows", ranks.numel(), "ms mean", float(np.mean(ts)), "min", float(np.min(ts)), "max", float(np.max(ts)))
PY
