"""GPU: the lift with cumulative_warp_features folded into its layout pass (fiery_lift_forward_warped; fiery/models/fiery.py:140-146,
fiery/utils/geometry.py:225-253) against (a) the oracle chain lift -> warp, (b) the unfused product chain (LiftSplat.forward ->
fiery_b200.warp.cumulative_warp_features), forward and backward, with and without a plan, through several passes of the scratch."""
import numpy as np
import pytest
import torch

from fiery_b200 import _lib
from fiery_b200.lift import LiftSplat
from fiery_b200.synthetic import CONFIGS, LiftConfig, make_calibration, make_egomotion, make_head
from fiery_b200.warp import cumulative_warp_features
from oracle import lift_oracle as O
from oracle import warp_oracle as W

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _case(cfg, b, s, seed):
    frames = b * s
    c = LiftConfig(**{**cfg.__dict__, "frames": frames})
    K, E = make_calibration(c, seed=seed)
    head = make_head(c, seed=seed)
    flow = make_egomotion(b, s, seed=seed)
    return c, torch.from_numpy(head), torch.from_numpy(K), torch.from_numpy(E), torch.from_numpy(flow)


def _extent(cfg):
    return float(cfg.x_bound[1]), float(cfg.y_bound[1])                   # Fiery.spatial_extent (fiery.py:31)


@pytest.mark.parametrize("name,b,s", [("cfg1_tiny", 2, 3), ("cfg1_tiny", 1, 2), ("cfg3_baseline", 1, 3)])
def test_fused_matches_oracle_chain_and_unfused_product(name, b, s):
    dev = torch.device("cuda:0")
    cfg, head, K, E, flow = _case(CONFIGS[name], b, s, seed=31)
    ext = _extent(cfg)
    lift = LiftSplat.from_config(cfg).to(dev)
    hd, Kd, Ed, fd = head.to(dev), K.to(dev), E.to(dev), flow.to(dev)
    fused = lift.forward_warped(hd, Kd, Ed, fd, ext)
    X, Y = cfg.bev_hw
    assert tuple(fused.shape) == (b, s, cfg.out_channels, X, Y) and fused.is_contiguous()
    unfused = cumulative_warp_features(lift(hd, Kd, Ed).unflatten(0, (b, s)), fd, mode="bilinear", spatial_extent=ext)
    scale = float(unfused.abs().max())
    assert float((fused - unfused).abs().max()) <= 5e-6 * scale            # same samples, same blend; only the lift's atomics differ
    exact = O.LiftOracle.from_config(cfg).lift_exact(head, K, E).unflatten(0, (b, s))
    want = W.cumulative_warp_features(exact.clone().float(), flow, mode="bilinear", spatial_extent=ext)
    assert float((fused.cpu() - want).abs().max()) <= TOL * float(want.abs().max())
    assert O.normwise_error(fused.cpu(), want) < TOL
    # the present frame passes through: it equals the plain lift of that frame to the lift's own run-to-run noise
    plain = lift(hd, Kd, Ed).unflatten(0, (b, s))
    assert float((fused[:, -1] - plain[:, -1]).abs().max()) <= 5e-6 * scale
    # the scratch is all-zero again (the next call would otherwise double-count): run twice, same answer
    again = lift.forward_warped(hd, Kd, Ed, fd, ext)
    assert float((again - fused).abs().max()) <= 5e-6 * scale


def test_fused_with_a_plan_and_in_several_passes():
    dev = torch.device("cuda:0")
    cfg, head, K, E, flow = _case(CONFIGS["cfg1_tiny"], 2, 3, seed=5)
    ext = _extent(cfg)
    lift = LiftSplat.from_config(cfg).to(dev)
    hd, Kd, Ed, fd = head.to(dev), K.to(dev), E.to(dev), flow.to(dev)
    ref = lift.forward_warped(hd, Kd, Ed, fd, ext)
    plan = lift.plan(Kd, Ed)
    with_plan = lift.forward_warped(hd, Kd, Ed, fd, ext, plan=plan)
    scale = float(ref.abs().max())
    assert float((with_plan - ref).abs().max()) <= 5e-6 * scale
    assert float((lift.forward_warped(hd, Kd, Ed, fd, ext, plan=plan) - ref).abs().max()) <= 5e-6 * scale    # the plan's marks survive
    lib = _lib.load()
    try:
        lib.fiery_lift_set_max_chunk_frames(2)                            # 6 frames in 3 passes over a 2-frame scratch
        chunked = lift.forward_warped(hd, Kd, Ed, fd, ext)
    finally:
        lib.fiery_lift_set_max_chunk_frames(0)
    assert float((chunked - ref).abs().max()) <= 5e-6 * scale


def test_single_frame_sequences_are_the_plain_lift():
    dev = torch.device("cuda:0")
    cfg, head, K, E, _ = _case(CONFIGS["cfg1_tiny"], 2, 1, seed=2)
    lift = LiftSplat.from_config(cfg).to(dev)
    hd, Kd, Ed = head.to(dev), K.to(dev), E.to(dev)
    out = lift.forward_warped(hd, Kd, Ed, torch.zeros(2, 1, 6, device=dev), (50.0, 50.0))
    plain = lift(hd, Kd, Ed)
    assert tuple(out.shape) == (2, 1, *plain.shape[1:])
    assert float((out[:, 0] - plain).abs().max()) <= 5e-6 * float(plain.abs().max())


def test_fused_gradient_matches_the_unfused_chain():
    dev = torch.device("cuda:0")
    cfg, head, K, E, flow = _case(CONFIGS["cfg1_tiny"], 2, 3, seed=11)
    ext = _extent(cfg)
    lift = LiftSplat.from_config(cfg).to(dev)
    Kd, Ed, fd = K.to(dev), E.to(dev), flow.to(dev)
    gout = torch.randn(2, 3, cfg.out_channels, *cfg.bev_hw, generator=torch.Generator().manual_seed(3)).to(dev)
    h1 = head.to(dev).requires_grad_(True)
    lift.forward_warped(h1, Kd, Ed, fd, ext).backward(gout)
    h2 = head.to(dev).requires_grad_(True)
    cumulative_warp_features(lift(h2, Kd, Ed).unflatten(0, (2, 3)), fd, mode="bilinear", spatial_extent=ext).backward(gout)
    assert O.normwise_error(h1.grad.cpu(), h2.grad.cpu()) < 1e-5
    # and against autograd through the oracle chain
    ho = head.clone().requires_grad_(True)
    bev = O.LiftOracle.from_config(cfg).lift(ho, K, E).unflatten(0, (2, 3))
    W.cumulative_warp_features(bev.clone(), flow, mode="bilinear", spatial_extent=ext).backward(gout.cpu())
    assert O.normwise_error(h1.grad.cpu(), ho.grad) < 1e-3


def test_c_abi_rejects_channels_last_and_null_maps():
    dev = torch.device("cuda:0")
    cfg, head, K, E, flow = _case(CONFIGS["cfg1_tiny"], 1, 2, seed=1)
    lift = LiftSplat.from_config(cfg).to(dev)
    lib = _lib.load()
    c = lift._constants(dev)
    desc = lift._desc(c, 2, cfg.n_cameras, torch.float32, _lib.CALIB_RAW, _lib.BEV_NHWC)
    out = torch.zeros(2, *cfg.bev_hw, cfg.out_channels, device=dev)
    th = torch.zeros(2, 6, device=dev)
    mk = torch.zeros(2, dtype=torch.uint8, device=dev)
    rc = lib.fiery_lift_forward_warped(desc, head.to(dev).data_ptr(), K.to(dev).data_ptr(), E.to(dev).data_ptr(), c["u"].data_ptr(),
                                       c["v"].data_ptr(), c["d"].data_ptr(), out.data_ptr(), 0, 0, th.data_ptr(), mk.data_ptr(), 0)
    assert rc != 0 and b"NCHW" in lib.fiery_last_error()
