"""Seeded synthetic inputs for the camera->BEV lift.

The reference has no dataset we can reach (nuScenes, ``fiery/data.py:350-363`` defines the tensor
contract), so every test and benchmark uses these generators.  Shapes follow the reference:

* head tensor  ``(B'*n, D+C, h, w)``  -- output of ``Encoder.depth_layer`` (``fiery/models/encoder.py:36,96``)
* intrinsics   ``(B', n, 3, 3)``, extrinsics ``(B', n, 4, 4)`` (camera -> ego), ``fiery/models/fiery.py:193-208``

Calibrations are nuScenes-like: 1600x900 images resized by 0.3 and top-cropped by 46 px
(``fiery/config.py:59-65``, ``fiery/utils/geometry.py:15-36``), six cameras at yaws
``[55, 0, -55, 110, 180, -110]`` degrees (order of ``IMAGE.NAMES``).  Everything is generated with numpy's
PCG64 so the same seed gives the same bytes on every machine.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Sequence, Tuple

import numpy as np

CAMERA_YAWS_DEG = (55.0, 0.0, -55.0, 110.0, 180.0, -110.0)


@dataclasses.dataclass(frozen=True)
class LiftConfig:
    """The subset of the reference config the lift reads (``fiery/config.py:59-78``)."""

    name: str
    n_cameras: int = 6
    final_dim: Tuple[int, int] = (224, 480)          # IMAGE.FINAL_DIM (H, W)
    downsample: int = 8                              # MODEL.ENCODER.DOWNSAMPLE
    out_channels: int = 64                           # MODEL.ENCODER.OUT_CHANNELS
    x_bound: Tuple[float, float, float] = (-50.0, 50.0, 0.5)
    y_bound: Tuple[float, float, float] = (-50.0, 50.0, 0.5)
    z_bound: Tuple[float, float, float] = (-10.0, 10.0, 20.0)
    d_bound: Tuple[float, float, float] = (2.0, 50.0, 1.0)
    frames: int = 1                                  # B' = batch x time receptive field
    use_depth_distribution: bool = True

    @property
    def feat_hw(self) -> Tuple[int, int]:
        return self.final_dim[0] // self.downsample, self.final_dim[1] // self.downsample

    @property
    def depth_bins(self) -> int:
        lo, hi, step = self.d_bound
        return int(math.ceil((hi - lo) / step))

    @property
    def bev_hw(self) -> Tuple[int, int]:
        return (int((self.x_bound[1] - self.x_bound[0]) / self.x_bound[2]),
                int((self.y_bound[1] - self.y_bound[0]) / self.y_bound[2]))

    @property
    def points_per_frame(self) -> int:
        h, w = self.feat_hw
        return self.n_cameras * self.depth_bins * h * w

    @property
    def head_channels(self) -> int:
        return self.out_channels + (self.depth_bins if self.use_depth_distribution else 0)

    def fwd_bytes_per_frame(self, head_itemsize: int = 4) -> int:
        """Algorithmic bytes per frame, forward: read head once + write BEV once."""
        h, w = self.feat_hw
        x, y = self.bev_hw
        return self.n_cameras * self.head_channels * h * w * head_itemsize + self.out_channels * x * y * 4

    def bwd_bytes_per_frame(self, head_itemsize: int = 4) -> int:
        """read grad-BEV + read head + write grad-head."""
        h, w = self.feat_hw
        x, y = self.bev_hw
        return self.out_channels * x * y * 4 + 2 * self.n_cameras * self.head_channels * h * w * head_itemsize


# BASELINE.json:configs, as concrete shapes.
CONFIGS = {
    # 1-cam 64x128 -> 50x50 BEV, the CPU correctness case
    "cfg1_tiny": LiftConfig("cfg1_tiny", n_cameras=1, final_dim=(64, 128),
                            x_bound=(-50.0, 50.0, 2.0), y_bound=(-50.0, 50.0, 2.0), frames=1),
    # literature/static_lss_setting.yml: 6-cam, 1 timestep, 200x200
    "cfg2_static_lss": LiftConfig("cfg2_static_lss", frames=1),
    "cfg2_static_lss_b8": LiftConfig("cfg2_static_lss_b8", frames=8),
    # baseline.yml: batch 3 x 3 past timesteps
    "cfg3_baseline": LiftConfig("cfg3_baseline", frames=9),
    # literature/pon_setting.yml: 400x200 @ 25 cm, batch 4 x 3
    "cfg4_pon": LiftConfig("cfg4_pon", x_bound=(-50.0, 50.0, 0.25), y_bound=(-25.0, 25.0, 0.25), frames=12),
    # not a reference YAML: cell sizes that are no powers of two, so s = (p - off) / res takes the true-division path
    # (fiery.py:236) instead of the exact scale -- a parity case, not a bench workload
    "cfg6_res_0p4_0p3": LiftConfig("cfg6_res_0p4_0p3", n_cameras=2, final_dim=(64, 160), x_bound=(-50.0, 50.0, 0.4),
                                   y_bound=(-30.0, 30.0, 0.3), frames=2),
}


def _rot_z(yaw: float) -> np.ndarray:
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _rodrigues(v: np.ndarray) -> np.ndarray:
    theta = float(np.linalg.norm(v))
    if theta < 1e-12:
        return np.eye(3)
    k = v / theta
    kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return np.eye(3) + math.sin(theta) * kx + (1.0 - math.cos(theta)) * (kx @ kx)


# camera axes (x right, y down, z forward) -> ego axes (x forward, y left, z up)
_CAM_TO_EGO = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


def make_calibration(cfg: LiftConfig, seed: int = 0, jitter_rad: float = 0.02,
                     per_frame_jitter: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Returns ``intrinsics (B', n, 3, 3)`` and ``extrinsics (B', n, 4, 4)`` as float32.

    ``jitter_rad=0`` gives the axis-aligned rig (front/back cameras at exactly 0/180 degrees) whose points land
    exactly on voxel boundaries -- the bit-exactness stress case of SURVEY.md section 4.
    """
    rng = np.random.default_rng(seed)
    H, W = cfg.final_dim
    scale = 0.3 * (W / 480.0)                       # keeps the nuScenes field of view at other image sizes
    yaws = [CAMERA_YAWS_DEG[i % len(CAMERA_YAWS_DEG)] for i in range(cfg.n_cameras)]
    if cfg.n_cameras == 1:
        yaws = [0.0]
    K = np.zeros((cfg.frames, cfg.n_cameras, 3, 3), dtype=np.float64)
    E = np.zeros((cfg.frames, cfg.n_cameras, 4, 4), dtype=np.float64)
    for f in range(cfg.frames):
        if f > 0 and not per_frame_jitter:
            K[f], E[f] = K[0], E[0]
            continue
        for i, yaw_deg in enumerate(yaws):
            focal = 1266.4 * scale * (1.0 + 0.01 * rng.standard_normal())
            rotvec = jitter_rad * rng.standard_normal(3)
            K[f, i] = [[focal, 0.0, 816.3 * scale], [0.0, focal, 491.5 * scale - 46.0 * (H / 224.0)], [0.0, 0.0, 1.0]]
            rz = _rot_z(math.radians(yaw_deg))
            R = rz @ _rodrigues(rotvec) @ _CAM_TO_EGO
            if jitter_rad == 0:
                R = np.round(R, 12)                  # cos(90 deg) etc. become exact zeros
            E[f, i, :3, :3] = R
            E[f, i, :3, 3] = rz @ np.array([1.5, 0.0, 0.0]) + np.array([0.0, 0.0, 1.5])
            E[f, i, 3, 3] = 1.0
    return K.astype(np.float32), E.astype(np.float32)


def make_head(cfg: LiftConfig, seed: int = 0, dtype=np.float32, scale: float = 1.0) -> np.ndarray:
    """``(B'*n, D+C, h, w)`` standard-normal logits/features (peaky-but-dense softmax)."""
    rng = np.random.default_rng(seed + 1000003)
    h, w = cfg.feat_hw
    x = rng.standard_normal((cfg.frames * cfg.n_cameras, cfg.head_channels, h, w), dtype=np.float32)
    return (x * np.float32(scale)).astype(dtype)


def make_grad_bev(cfg: LiftConfig, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed + 2000003)
    x, y = cfg.bev_hw
    return rng.standard_normal((cfg.frames, cfg.out_channels, x, y), dtype=np.float32)


def shard_frames(n_frames: int, world_size: int, rank: int) -> range:
    """Contiguous batch x time shard owned by ``rank`` (DistributedSampler-style split of independent frames,
    SURVEY.md section 8e).  The lift has no cross-frame dependency (``fiery/models/fiery.py:231``)."""
    base, rem = divmod(n_frames, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def make_egomotion(batch: int, steps: int, seed: int = 0) -> np.ndarray:
    """``future_egomotion (b, s, 6)`` as fiery/data.py:350-363 provides it: per-step 6-DoF ego motion (tx, ty, tz, rx, ry, rz).
    Driving-like: ~0.5 s between frames at 5-15 m/s forward, small lateral slip, yaw rate up to ~0.2 rad/step."""
    rng = np.random.default_rng(seed + 3000017)
    v = np.zeros((batch, steps, 6), dtype=np.float32)
    v[..., 0] = rng.uniform(2.5, 7.5, (batch, steps))
    v[..., 1] = rng.normal(0.0, 0.2, (batch, steps))
    v[..., 2] = rng.normal(0.0, 0.02, (batch, steps))
    v[..., 3] = rng.normal(0.0, 0.005, (batch, steps))
    v[..., 4] = rng.normal(0.0, 0.005, (batch, steps))
    v[..., 5] = rng.normal(0.0, 0.08, (batch, steps))
    return v
