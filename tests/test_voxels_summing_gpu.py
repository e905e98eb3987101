"""GPU parity of the VoxelsSumming drop-in (fiery/utils/geometry.py:283-314) through the C ABI."""
import numpy as np
import pytest
import torch

from fiery_b200.geometry import VoxelsSumming
from oracle import lift_oracle as O

pytestmark = pytest.mark.gpu
PATTERNS = ["singletons", "one_voxel", "long_runs", "first_last_boundaries", "random_runs", "empty", "single_row"]


@pytest.mark.parametrize("name", PATTERNS)
def test_patterns_match_reference(golden_vs, name):
    dev = torch.device("cuda:0")
    feats = torch.from_numpy(golden_vs[f"{name}__feats"]).to(dev).requires_grad_(True)
    coords = torch.from_numpy(golden_vs[f"{name}__coords"]).to(dev)
    ranks = torch.from_numpy(golden_vs[f"{name}__ranks"]).to(dev)
    sums, kept = VoxelsSumming.apply(feats, coords, ranks)
    ref = torch.from_numpy(golden_vs[f"{name}__ref_sum"])
    assert tuple(sums.shape) == tuple(ref.shape)
    assert torch.equal(kept.cpu(), torch.from_numpy(golden_vs[f"{name}__ref_coords"]))          # integers: bit-exact
    assert not kept.requires_grad
    if ranks.numel():
        assert torch.allclose(sums.detach().cpu(), ref, rtol=1e-4, atol=1e-5)
        exact = O.direct_segment_sum(feats.detach().cpu(), ranks.cpu())
        assert torch.allclose(sums.detach().cpu().double(), exact, rtol=1e-5, atol=1e-6)
        sums.backward(torch.from_numpy(golden_vs[f"{name}__gout"]).to(dev))
        assert np.array_equal(feats.grad.cpu().numpy(), golden_vs[f"{name}__ref_grad"])         # pure gather: bit-exact


def test_full_size_sorted_volume():
    """Nm ~ 450k rows x 64 channels, runs of very different length (the shape of one cfg-2 frame)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    n_vox = 12927
    lengths = torch.randint(1, 70, (n_vox,), generator=g)
    lengths[::97] = 420
    ranks = torch.repeat_interleave(torch.arange(n_vox) * 3, lengths)
    Nm = ranks.numel()
    feats = torch.randn(Nm, 64, generator=g)
    coords = torch.stack([ranks // 200, ranks % 200, torch.zeros_like(ranks)], 1)
    x = feats.to(dev).requires_grad_(True)
    sums, kept = VoxelsSumming.apply(x, coords.to(dev), ranks.to(dev))
    assert sums.shape == (n_vox, 64)
    exact = O.direct_segment_sum(feats, ranks)
    ref, ref_kept = O.CumsumSegmentSum.apply(feats, coords, ranks)
    assert torch.equal(kept.cpu(), ref_kept)
    e_ours, e_ref = O.normwise_error(sums.cpu(), exact), O.normwise_error(ref, exact)
    assert e_ours < 1e-6 and e_ours <= e_ref
    assert O.normwise_error(sums.cpu(), ref) < 1e-4
    gout = torch.randn(n_vox, 64, generator=g)
    sums.backward(gout.to(dev))
    seg = torch.repeat_interleave(torch.arange(n_vox), lengths)
    assert torch.equal(x.grad.cpu(), gout[seg])


def test_non_contiguous_rows():
    dev = torch.device("cuda:0")
    base = torch.randn(100, 24, device=dev)
    x = base[:, :16]                                     # row stride 24
    ranks = torch.sort(torch.randint(0, 9, (100,), device=dev)).values
    coords = torch.stack([ranks, ranks, ranks * 0], 1)
    sums, _ = VoxelsSumming.apply(x, coords, ranks)
    exact = O.direct_segment_sum(x.cpu(), ranks.cpu())
    assert torch.allclose(sums.cpu().double(), exact, rtol=1e-5, atol=1e-6)
