#!/usr/bin/env python
"""Target for compute-sanitizer (memcheck / racecheck): every kernel added late in round 2, once, at small sizes.
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fiery_b200.depth_layer import depth_layer_forward
from fiery_b200.lift import LiftSplat
from fiery_b200.synthetic import CONFIGS, LiftConfig, make_calibration, make_egomotion, make_head
from fiery_b200.warp import cumulative_warp_features

dev = torch.device("cuda:0")
for dt in (torch.float16, torch.float32):                       # depth_layer_kernel<2> / <4>: ragged last tile, 100 of 128 channels
    feat = torch.randn(3, 128, 9, 24, device=dev).to(dt)        # 216 pixels: a 16-byte row pitch in both dtypes, ragged last tile
    depth_layer_forward(feat, torch.randn(100, 128, 1, 1, device=dev), torch.randn(100, device=dev))
cfg = LiftConfig(**{**CONFIGS["cfg1_tiny"].__dict__, "frames": 4})
K, E = (torch.from_numpy(a).to(dev) for a in make_calibration(cfg, seed=1))
head = torch.from_numpy(make_head(cfg, seed=1)).to(dev).requires_grad_(True)
flow = torch.from_numpy(make_egomotion(2, 2, seed=0)).to(dev)
lift = LiftSplat.from_config(cfg).to(dev)
out = lift.forward_warped(head, K, E, flow, (float(cfg.x_bound[1]), float(cfg.y_bound[1])))   # finalize_warp + clear_touched (+ plan)
out.sum().backward()                                            # warp_backward_gather + nchw_to_nhwc + lift_backward
x = torch.randn(2, 3, 5, 24, 40, device=dev, requires_grad=True)
f3 = torch.from_numpy(make_egomotion(2, 3, seed=3)).to(dev)
cumulative_warp_features(x, f3, mode="bilinear", spatial_extent=(50.0, 50.0)).sum().backward()
cumulative_warp_features(x.detach(), f3, mode="nearest", spatial_extent=(50.0, 50.0))
torch.cuda.synchronize()
print("sanitize target done")
