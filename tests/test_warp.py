"""cumulative_warp_features / warp_features (fiery/utils/geometry.py:181-253).
CPU: the oracle reproduces the reference's recorded outputs.  GPU: the CUDA path (C ABI) matches the oracle."""
import os

import numpy as np
import pytest
import torch

from fiery_b200.synthetic import make_egomotion
from oracle import warp_oracle as W

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp.npz")
TOL = 1e-4       # north_star: fp32 features within 1e-4 relative


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("tag", ["small", "rect"])
def test_oracle_matches_reference(golden, tag):
    x = torch.from_numpy(golden[f"{tag}__x"]).requires_grad_(True)
    flow = torch.from_numpy(golden[f"{tag}__flow"])
    extent = tuple(float(v) for v in golden[f"{tag}__extent"])
    out = W.cumulative_warp_features(x.clone(), flow, mode="bilinear", spatial_extent=extent)
    assert torch.allclose(out.detach(), torch.from_numpy(golden[f"{tag}__ref"]), rtol=1e-6, atol=1e-6)
    out.backward(torch.from_numpy(golden[f"{tag}__gout"]))
    assert torch.allclose(x.grad, torch.from_numpy(golden[f"{tag}__grad"]), rtol=1e-5, atol=1e-6)
    assert torch.equal(out[:, -1].detach(), x[:, -1].detach())            # present frame untouched (geometry.py:243)


def test_oracle_single_frame_is_identity():
    x = torch.randn(2, 1, 3, 8, 8)
    assert W.cumulative_warp_features(x, torch.zeros(2, 1, 6), mode="bilinear", spatial_extent=(50.0, 50.0)) is x


def _gpu_case(golden, tag):
    b, t, c, h, w = (int(v) for v in golden[f"{tag}__shape"])
    flow = torch.from_numpy(golden[f"{tag}__flow"])
    extent = tuple(float(v) for v in golden[f"{tag}__extent"])
    if f"{tag}__x" in golden.files:
        x = torch.from_numpy(golden[f"{tag}__x"])
    else:
        torch.manual_seed(0)                                            # same random stream as oracle/gen_golden_warp.py:
        xs = {}                                                         # per case x = randn(shape), then gout = randn(shape)
        for tg, shp in (("small", (2, 3, 5, 12, 16)), ("rect", (1, 4, 3, 20, 10)), ("bev", (1, 3, 8, 200, 200))):
            xs[tg] = torch.randn(*shp)
            torch.randn(*shp)
        x = xs[tag]
    return x, flow, extent


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["small", "rect", "bev"])
def test_gpu_matches_oracle_and_reference(golden, tag):
    from fiery_b200.warp import cumulative_warp_features, warp_features
    dev = torch.device("cuda:0")
    x, flow, extent = _gpu_case(golden, tag)
    xd = x.to(dev).requires_grad_(True)
    out = cumulative_warp_features(xd, flow.to(dev), mode="bilinear", spatial_extent=extent)
    assert out.shape == x.shape and out.dtype == torch.float32
    xo = x.clone().requires_grad_(True)
    ref = W.cumulative_warp_features(xo.clone(), flow, mode="bilinear", spatial_extent=extent)
    scale = float(ref.abs().max())
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= TOL * scale
    assert torch.equal(out[:, -1].detach().cpu(), x[:, -1])                # present frame bit-exact
    gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    out.backward(gout.to(dev))
    ref.backward(gout)
    gscale = float(xo.grad.abs().max())
    assert float((xd.grad.cpu() - xo.grad).abs().max()) <= TOL * gscale
    if f"{tag}__ref_at_pick" in golden.files:                              # the reference's own recorded samples
        pick = golden[f"{tag}__pick"]
        assert np.abs(out.detach().cpu().flatten()[pick].numpy() - golden[f"{tag}__ref_at_pick"]).max() <= TOL * scale
    # single-map entry point, both sampling modes
    for mode in ("bilinear", "nearest"):
        one = warp_features(x[:, 0].to(dev), flow[:, 0].to(dev), mode=mode, spatial_extent=extent).cpu()
        ref1 = W.warp_features(x[:, 0], flow[:, 0], mode=mode, spatial_extent=extent)
        if mode == "bilinear":
            assert float((one - ref1).abs().max()) <= TOL * scale
        else:       # nearest: a 1-ulp coordinate difference may flip a pick exactly between two pixels; allow a handful
            assert float(((one - ref1).abs() > TOL * scale).float().mean()) < 1e-3


@pytest.mark.gpu
def test_gpu_full_size_properties():
    """Full benchmark size (3 x 3 frames of 64 x 200 x 200): zero ego motion is the identity up to fp32 rounding of the
    sample positions; the warp is linear in x."""
    from fiery_b200.warp import cumulative_warp_features
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 3, 64, 200, 200, generator=g).to(dev)
    y = torch.randn(3, 3, 64, 200, 200, generator=g).to(dev)
    still = cumulative_warp_features(x, torch.zeros(3, 3, 6, device=dev), mode="bilinear", spatial_extent=(50.0, 50.0))
    assert float((still - x).abs().max()) < 1e-3
    flow = torch.from_numpy(make_egomotion(3, 3, seed=2)).to(dev)
    a = cumulative_warp_features(x, flow, mode="bilinear", spatial_extent=(50.0, 50.0))
    b = cumulative_warp_features(y, flow, mode="bilinear", spatial_extent=(50.0, 50.0))
    c = cumulative_warp_features(0.5 * x + y, flow, mode="bilinear", spatial_extent=(50.0, 50.0))
    assert float((0.5 * a + b - c).abs().max()) < 1e-4 * float(c.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("T", [2, 3, 5])
def test_gpu_pose_algebra_matches_oracle(T):
    """fiery_warp_theta (pose_vec2mat, running product, mat2pose_vec, theta) against the oracle's torch-CPU restatement of
    geometry.py:241-251 and :197-219; fp32, different summation order only."""
    from fiery_b200.warp import _device_theta
    dev = torch.device("cuda:0")
    flow = torch.from_numpy(make_egomotion(4, T, seed=7 + T))
    flow[:, :, 3:5] += 0.01 * torch.randn(4, T, 2, generator=torch.Generator().manual_seed(T))     # roll / pitch too
    ext = (50.0, 25.0)
    want = W.cumulative_warp_thetas(flow, ext)
    theta, mask = _device_theta(flow.to(dev), ext, cumulative=True)
    theta, mask = theta.cpu().view(4, T, 2, 3), mask.cpu().view(4, T)
    assert mask[:, -1].eq(1).all() and mask[:, :-1].eq(0).all()
    for t in range(T - 1):
        np.testing.assert_allclose(theta[:, t].numpy(), want[t].numpy(), rtol=0, atol=2e-6)
    plain, none = _device_theta(flow[:, 0].to(dev), ext, cumulative=False)
    assert none is None
    np.testing.assert_allclose(plain.cpu().numpy(), W.warp_theta(flow[:, 0], ext).numpy(), rtol=0, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("nearest", [0, 1])
def test_gpu_backward_of_arbitrary_affine_maps(nearest):
    """fiery_warp_features_backward through the C ABI for maps that are NOT the rotations warp_features builds: the gather adjoint
    covers near-rigid maps, everything else (scaling, shear, singular, zero, non-finite) takes the scatter inside the same call;
    grad_x is overwritten either way (it is handed in full of garbage).  Reference: autograd of affine_grid + grid_sample."""
    import torch.nn.functional as F
    from fiery_b200 import _lib
    from fiery_b200.geometry import _stream_ptr
    dev = torch.device("cuda:0")
    lib = _lib.load()
    C, H, W = 5, 24, 40
    th = torch.tensor([
        [[0.96, -0.28, 0.10], [0.28, 0.96, -0.20]],        # rotation (gather)
        [[1.00, 0.00, 0.00], [0.00, 1.00, 0.00]],          # identity (gather)
        [[0.50, 0.00, 0.30], [0.00, 0.45, 0.00]],          # magnification: determinant < 1/4 (scatter)
        [[3.00, 0.50, 0.00], [0.20, 2.50, 0.10]],          # minification: many outputs per source... still gather-sized
        [[0.00, 0.00, 0.20], [0.00, 0.00, -0.30]],         # singular: every output samples one point (scatter)
        [[1.00, 0.90, 0.00], [0.00, 1.00, 0.00]],          # shear (gather)
        [[-0.60, 0.80, 0.00], [-0.80, -0.60, 0.50]],       # rotation by more than 90 degrees (gather)
    ], dtype=torch.float32, device=dev)
    n = th.shape[0]
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(n, C, H, W, generator=gen).to(dev).requires_grad_(True)
    gout = torch.randn(n, C, H, W, generator=gen).to(dev)
    grid = F.affine_grid(th, (n, C, H, W), align_corners=False)
    F.grid_sample(x, grid, mode="nearest" if nearest else "bilinear", padding_mode="zeros", align_corners=False).backward(gout)
    got = torch.full((n, C, H, W), float("nan"), device=dev)
    _lib.check(lib.fiery_warp_features_backward(n, C, H, W, gout.data_ptr(), C * H * W, th.data_ptr(), 0, got.data_ptr(), C * H * W,
                                                nearest, _stream_ptr(dev)), "warp backward")
    assert torch.isfinite(got).all()
    scale = float(x.grad.abs().max())
    if nearest:        # a pick exactly between two pixels may flip with 1-ulp coordinate differences
        assert float(((got - x.grad).abs() > 1e-4 * scale).float().mean()) < 5e-3
    else:
        assert float((got - x.grad).abs().max()) <= 1e-4 * scale
    # a non-finite map gives a zero gradient (all of its samples are out of range), and does not disturb its neighbours
    th2 = th.clone()
    th2[1, 0, 0] = float("nan")
    got2 = torch.full((n, C, H, W), float("nan"), device=dev)
    _lib.check(lib.fiery_warp_features_backward(n, C, H, W, gout.data_ptr(), C * H * W, th2.data_ptr(), 0, got2.data_ptr(), C * H * W,
                                                nearest, _stream_ptr(dev)), "warp backward")
    assert float(got2[1].abs().max()) == 0.0 and torch.equal(got2[0], got[0]) and torch.equal(got2[3], got[3])


@pytest.mark.gpu
def test_gpu_backward_is_deterministic_at_full_size():
    from fiery_b200.warp import cumulative_warp_features
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 3, 64, 200, 200, generator=g).to(dev).requires_grad_(True)
    gout = torch.randn(2, 3, 64, 200, 200, generator=g).to(dev)
    flow = torch.from_numpy(make_egomotion(2, 3, seed=4)).to(dev)
    grads = []
    for _ in range(2):
        x.grad = None
        cumulative_warp_features(x, flow, mode="bilinear", spatial_extent=(50.0, 50.0)).backward(gout)
        grads.append(x.grad.clone())
    assert torch.equal(grads[0], grads[1])                                  # gather adjoint: no atomics on this path
    assert torch.equal(grads[0][:, -1], gout[:, -1])                        # the present frame's gradient passes through
    # adjoint identity <W x, g> == <x, W^T g>
    with torch.no_grad():
        y = cumulative_warp_features(x.detach(), flow, mode="bilinear", spatial_extent=(50.0, 50.0))
    lhs, rhs = float((y.double() * gout.double()).sum()), float((x.detach().double() * grads[0].double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
