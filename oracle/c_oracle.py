"""ctypes access to the plain-C oracle (oracle/lift_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liblift_oracle.so")


def load():
    if not os.path.exists(_LIB):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return ctypes.CDLL(_LIB)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def voxel_indices(u, v, depth, combined, translation, offset, res, dim):
    """combined (B',n,3,3), translation (B',n,3) float32 -> idx (B',N,3) int64, keep (B',N) bool."""
    lib = load()
    B, n = combined.shape[:2]
    D, h, w = depth.size, v.size, u.size
    N = n * D * h * w
    idx = np.empty((B, N, 3), dtype=np.int64)
    keep = np.empty((B, N), dtype=np.uint8)
    args = [np.ascontiguousarray(x, dtype=np.float32) for x in (u, v, depth, combined, translation, offset, res)]
    dim = np.ascontiguousarray(dim, dtype=np.int64)
    lib.oracle_voxel_indices(B, n, D, h, w, *[_p(a) for a in args], _p(dim), _p(idx), _p(keep))
    return idx, keep.astype(bool)


def pool_exact(prob, ctx, idx, keep, n_cams, X, Y):
    """prob (B'n,D,h,w), ctx (B'n,C,h,w) float64 -> bev (B',C,X,Y) float64."""
    lib = load()
    Bn, D, h, w = prob.shape
    C = ctx.shape[1]
    B = Bn // n_cams
    bev = np.zeros((B, C, X, Y), dtype=np.float64)
    prob = np.ascontiguousarray(prob, dtype=np.float64)
    ctx = np.ascontiguousarray(ctx, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    keep8 = np.ascontiguousarray(keep, dtype=np.uint8)
    lib.oracle_pool_exact(B, n_cams, D, C, h, w, _p(prob), _p(ctx), _p(idx), _p(keep8), X, Y, _p(bev))
    return bev
