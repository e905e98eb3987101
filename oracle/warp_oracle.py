"""CPU oracle for the step right after the lift: ``cumulative_warp_features``.  TEST INFRASTRUCTURE ONLY.

Restates fiery/utils/geometry.py:110-253 (euler2mat, pose_vec2mat, mat2pose_vec, warp_features,
cumulative_warp_features) with the same torch primitives in the same order -- results are defined by torch's
``affine_grid`` / ``grid_sample`` (align_corners=False, zero padding) and small batched matmuls.  Call site:
fiery/models/fiery.py:143-146 (``mode='bilinear'``, ``spatial_extent=(X_BOUND[1], Y_BOUND[1])``).
Pinned against the real reference functions by oracle/gen_golden.py (tests/golden/warp.npz).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def euler_to_matrix(angle: torch.Tensor) -> torch.Tensor:
    """(..., 3) Euler angles -> (..., 3, 3) = Rx @ Ry @ Rz; geometry.py:110-142."""
    shape = angle.shape
    a = angle.reshape(-1, 3)
    rx, ry, rz = a[:, 0], a[:, 1], a[:, 2]
    zero, one = torch.zeros_like(rz), torch.ones_like(rz)
    cz, sz = torch.cos(rz), torch.sin(rz)
    zmat = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], dim=1).view(-1, 3, 3)
    cy, sy = torch.cos(ry), torch.sin(ry)
    ymat = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], dim=1).view(-1, 3, 3)
    cx, sx = torch.cos(rx), torch.sin(rx)
    xmat = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], dim=1).view(-1, 3, 3)
    return xmat.bmm(ymat).bmm(zmat).view(*shape[:-1], 3, 3)                       # geometry.py:140


def pose_vector_to_matrix(vec: torch.Tensor) -> torch.Tensor:
    """(..., 6) (tx,ty,tz,rx,ry,rz) -> (..., 4, 4); geometry.py:145-160."""
    rot = euler_to_matrix(vec[..., 3:].contiguous())
    top = torch.cat([rot, vec[..., :3].unsqueeze(-1)], dim=-1)
    mat = F.pad(top, [0, 0, 0, 1], value=0)
    mat[..., 3, 3] = 1.0
    return mat


def matrix_to_pose_vector(m: torch.Tensor) -> torch.Tensor:
    """(..., 4, 4) -> (..., 6); geometry.py:82-107."""
    rotx = torch.atan2(-m[..., 1, 2], m[..., 2, 2])
    cosy = torch.sqrt(m[..., 1, 2] ** 2 + m[..., 2, 2] ** 2)
    roty = torch.atan2(m[..., 0, 2], cosy)
    rotz = torch.atan2(-m[..., 0, 1], m[..., 0, 0])
    return torch.cat((m[..., :3, 3], torch.stack((rotx, roty, rotz), dim=-1)), dim=-1)


def warp_theta(flow: torch.Tensor, spatial_extent) -> torch.Tensor:
    """The (b, 2, 3) affine map warp_features builds from a 6-DoF vector (z-rotation + xy translation);
    geometry.py:197-219."""
    angle = flow[:, 5].clone()
    translation = flow[:, :2].clone()
    translation[:, 0] /= spatial_extent[0]
    translation[:, 1] /= spatial_extent[1]
    translation[:, 0] *= -1
    c, s = torch.cos(angle), torch.sin(angle)
    return torch.stack([c, -s, translation[:, 1], s, c, translation[:, 0]], dim=-1).view(-1, 2, 3)


def warp_features(x: torch.Tensor, flow: torch.Tensor, mode: str = "nearest", spatial_extent=None) -> torch.Tensor:
    """geometry.py:181-222."""
    if flow is None:
        return x
    theta = warp_theta(flow, spatial_extent)
    grid = F.affine_grid(theta, size=x.shape, align_corners=False)                  # geometry.py:219
    return F.grid_sample(x, grid.float(), mode=mode, padding_mode="zeros", align_corners=False)


def cumulative_warp_thetas(flow: torch.Tensor, spatial_extent):
    """The per-timestep affine maps cumulative_warp_features applies: list over t = T-2 .. 0 of (b, 2, 3);
    geometry.py:241-251 (cum_flow starts at flow[:, -2] and is left-multiplied by flow[:, t-1])."""
    mats = pose_vector_to_matrix(flow)
    T = flow.shape[1]
    out = {}
    cum = mats[:, -2]
    for t in reversed(range(T - 1)):
        out[t] = warp_theta(matrix_to_pose_vector(cum), spatial_extent)
        cum = mats[:, t - 1] @ cum
    return out


def cumulative_warp_features(x: torch.Tensor, flow: torch.Tensor, mode: str = "nearest", spatial_extent=None) -> torch.Tensor:
    """(b, t, c, h, w) features, (b, t, 6) egomotion -> every past frame warped into the present frame; geometry.py:225-253."""
    T = x.shape[1]
    if T == 1:
        return x
    mats = pose_vector_to_matrix(flow)
    out = [x[:, -1]]
    cum = mats[:, -2]
    for t in reversed(range(T - 1)):
        out.append(warp_features(x[:, t], matrix_to_pose_vector(cum), mode=mode, spatial_extent=spatial_extent))
        cum = mats[:, t - 1] @ cum
    return torch.stack(out[::-1], 1)
