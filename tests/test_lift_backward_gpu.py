"""GPU parity of the lift backward (grad w.r.t. the head tensor) through the C ABI, against autograd through the oracle's
restatement of the reference (softmax/outer product encoder.py:99-100, VoxelsSumming.backward geometry.py:305-314)."""
import numpy as np
import pytest
import torch

from fiery_b200.lift import LiftSplat
from fiery_b200.synthetic import CONFIGS, LiftConfig, make_calibration, make_grad_bev, make_head
from oracle import lift_oracle as O
from tests._cases import GOLDEN_CASES, build_case, case_id, golden_tag

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _oracle_grad(cfg, head, K, E, gout, comb=None, exact=False):
    oracle = O.LiftOracle.from_config(cfg)
    h = head.clone().double().requires_grad_(True) if exact else head.clone().requires_grad_(True)
    if exact:
        # fp64 autograd through a direct scatter-add: ground truth for the gradient
        n = K.shape[1]
        idx, keep = oracle.point_indices(K, E, comb)
        vol = O.depth_context_volume(h, n, oracle.D, oracle.C, oracle.use_depth_distribution)
        X, Y, _ = (int(d) for d in oracle.dimension)
        total = 0.0
        for b in range(vol.shape[0]):
            feats = vol[b].reshape(-1, oracle.C)[keep[b]]
            cell = idx[b][keep[b]]
            bev = torch.zeros(X * Y, oracle.C, dtype=torch.float64).index_add(0, cell[:, 0] * Y + cell[:, 1], feats)
            total = total + (bev.view(X, Y, oracle.C).permute(2, 0, 1) * gout[b].double()).sum()
        total.backward()
    else:
        oracle.lift(h, K, E, combined=comb).backward(gout)
    return h.grad


@pytest.mark.parametrize("grad_layout", ["contiguous", "channels_last"])
@pytest.mark.parametrize("case", GOLDEN_CASES, ids=case_id)
def test_backward_matches_oracle(golden_lift, case, grad_layout):
    cfg, K, E, head, gout = build_case(case)
    tag = golden_tag(case)
    dev = torch.device("cuda:0")
    lift = LiftSplat.from_config(cfg).to(dev)
    hd = head.to(dev).requires_grad_(True)
    bev = lift(hd, K.to(dev), E.to(dev))
    g = gout.to(dev)
    if grad_layout == "channels_last":
        g = g.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    bev.backward(g)
    got = hd.grad.cpu()
    assert got.shape == head.shape and got.dtype == head.dtype
    comb = torch.from_numpy(golden_lift[f"{tag}__combined"])
    ref = _oracle_grad(cfg, head, K, E, gout, comb)
    exact = _oracle_grad(cfg, head, K, E, gout, comb, exact=True)
    assert O.normwise_error(got, ref) < TOL and O.max_abs_scaled_error(got, ref) < TOL
    assert O.normwise_error(got, exact) < TOL and O.max_abs_scaled_error(got, exact) < TOL
    assert O.normwise_error(got, exact) <= max(O.normwise_error(ref, exact), 2e-6)
    # the reference's own recorded gradient samples
    pick = golden_lift[f"{tag}__grad_pick"]
    rec = golden_lift[f"{tag}__grad_ref_at_pick"]
    assert np.abs(got.reshape(-1)[pick].numpy() - rec).max() <= TOL * float(np.abs(rec).max())
    assert abs(float(got.double().norm()) / float(golden_lift[f"{tag}__grad_norm"][0]) - 1.0) < 1e-4


def test_backward_uniform_depth_and_masked_points():
    base = CONFIGS["cfg1_tiny"]
    cfg = LiftConfig(**{**base.__dict__, "use_depth_distribution": False})
    dev = torch.device("cuda:0")
    K, E = make_calibration(cfg, seed=2)
    K, E = torch.from_numpy(K), torch.from_numpy(E)
    head = torch.from_numpy(make_head(cfg, seed=2))
    gout = torch.from_numpy(make_grad_bev(cfg, seed=2))
    lift = LiftSplat.from_config(cfg).to(dev)
    hd = head.to(dev).requires_grad_(True)
    lift(hd, K.to(dev), E.to(dev)).backward(gout.to(dev))
    exact = _oracle_grad(cfg, head, K, E, gout, exact=True)
    assert O.normwise_error(hd.grad.cpu(), exact) < TOL


def test_gradcheck_directional_full_size():
    """Size-independent property at full size: <grad_head, dh> == d/dt sum(lift(head + t*dh) * gout) (the lift is
    smooth in the head tensor; indices do not depend on it)."""
    cfg = LiftConfig(**{**CONFIGS["cfg3_baseline"].__dict__, "frames": 3})
    dev = torch.device("cuda:0")
    K, E = make_calibration(cfg, seed=9)
    K, E = torch.from_numpy(K).to(dev), torch.from_numpy(E).to(dev)
    head = torch.from_numpy(make_head(cfg, seed=9)).to(dev)
    dh = torch.from_numpy(make_head(cfg, seed=10)).to(dev)
    gout = torch.from_numpy(make_grad_bev(cfg, seed=9)).to(dev)
    lift = LiftSplat.from_config(cfg).to(dev)
    hd = head.clone().requires_grad_(True)
    lift(hd, K, E).backward(gout)
    analytic = float((hd.grad.double() * dh.double()).sum())
    eps = 1e-2
    with torch.no_grad():
        fp = float((lift(head + eps * dh, K, E).double() * gout.double()).sum())
        fm = float((lift(head - eps * dh, K, E).double() * gout.double()).sum())
    numeric = (fp - fm) / (2 * eps)
    assert abs(analytic - numeric) <= 2e-3 * max(abs(analytic), abs(numeric))
