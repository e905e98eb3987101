#!/usr/bin/env python
"""A/B of the backward tile kernel on an experiment build (FIERY_NVCC_EXTRA=-DFIERY_COLS_AB): FIERY_BWD_EARLY=0/1, channel-last and
NCHW gradient, with the forward's plan; each variant is checked against the first."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from fiery_b200.lift import LiftSplat
from fiery_b200.synthetic import CONFIGS, make_calibration, make_grad_bev, make_head

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_static_lss_b8"
cfg = CONFIGS[wl]
dev = torch.device("cuda:0")
K, E = make_calibration(cfg, seed=100)
head = torch.from_numpy(make_head(cfg, seed=100)).to(dev)
K_d, E_d = torch.from_numpy(K).to(dev), torch.from_numpy(E).to(dev)
g = torch.from_numpy(make_grad_bev(cfg, seed=100)).to(dev)
g_cl = g.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
lift = LiftSplat.from_config(cfg).to(dev)
plan = lift.plan(K_d, E_d)
flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)


def timed(fn, n=20):
    ts = []
    for _ in range(n):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)) * 1e3, float(np.min(ts)) * 1e3


ref = None
for early in ("0", "1"):
    os.environ["FIERY_BWD_EARLY"] = early
    for name, grad in (("channels_last", g_cl), ("nchw", g)):
        out = lift._launch_backward(head, K_d, E_d, grad, plan=plan)
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        err = float((out - ref).norm() / ref.norm())
        for _ in range(3):
            lift._launch_backward(head, K_d, E_d, grad, plan=plan)
        mean, mn = timed(lambda: lift._launch_backward(head, K_d, E_d, grad, plan=plan))
        print("bwd", wl, "early", early, name, {"us_mean": mean, "us_min": mn, "rel_err": err}, flush=True)
