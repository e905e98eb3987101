"""Data-parallel training step around the fused lift.

The reference trains through PyTorch Lightning: ``pl.Trainer(accelerator='ddp', precision=cfg.PRECISION, sync_batchnorm=True,
gradient_clip_val=cfg.GRAD_NORM_CLIP)`` (train.py:33-46), ``TrainingModule.training_step`` = forward + uncertainty-weighted losses
(fiery/trainer.py:69-120,200-208) and ``Adam(lr=3e-4, weight_decay=1e-7)`` (trainer.py:254-260, config.py:121-123).  This module is
that step without Lightning, one process per GPU:

  * ``FlatGradBucket`` -- every parameter's gradient is a view of ONE flat fp32 buffer, so the data-parallel exchange is a single
    ``all_reduce`` per step (NCCL over NVLink/NVSwitch; DDP in the reference issues one per 25 MB bucket);
  * ``LiftTrainModel`` -- ``Encoder.get_features`` stand-in -> the reference's ``depth_layer`` (1x1 conv, encoder.py:36,96) -> the
    fused CUDA lift (fiery_b200.lift, forward AND backward, sharing one geometry plan) -> a small BEV head.  The image backbone
    (EfficientNet) and the temporal model / decoder are OUT of scope here: the stand-ins only give the lift a
    producer and a consumer with parameters, so that the step has real autograd, AMP, clipping, optimiser and all-reduce around it;
  * ``synthetic_batch`` -- tensors with the keys, shapes and dtypes of ``FuturePredictionDataset.__getitem__``
    (fiery/data.py:343-363) batched by the loader.

No CPU path: the lift raises on CPU tensors.  ``FlatGradBucket`` itself is device-agnostic (tested with gloo on CPU).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .depth_layer import DepthLayer
from .lift import LiftSplat, pack_sequence_dim, unpack_sequence_dim
from .synthetic import LiftConfig, make_calibration, make_egomotion, shard_frames

# fiery/config.py:121-123
LR, WEIGHT_DECAY, GRAD_NORM_CLIP = 3e-4, 1e-7, 5.0


class FlatGradBucket:
    """Gradients of ``params`` as views of one flat fp32 buffer; ``all_reduce_mean`` is the step's only collective."""

    def __init__(self, params: Iterable[nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatGradBucket expects fp32 parameters on one device (AMP keeps master weights in fp32)")
            p.grad = self.flat[off:off + p.numel()].view_as(p)        # autograd accumulates into the view in place
            off += p.numel()

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce_mean(self, group=None) -> None:
        """Average over the data-parallel group: ONE collective call on the flat buffer."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


class StandInEncoder(nn.Module):
    """Same interface as the reference ``Encoder`` (fiery/models/encoder.py:9-104): ``get_features`` (image -> 128-channel map at
    1/8 resolution; a three-layer strided conv stack instead of EfficientNet, which is not available offline) and the reference's
    ``depth_layer`` 1x1 conv to D + C channels (encoder.py:36).  ``forward`` returns the head tensor the lift consumes."""

    def __init__(self, depth_bins: int, out_channels: int, use_depth_distribution: bool = True, width: int = 128):
        super().__init__()
        self.D, self.C, self.use_depth_distribution = depth_bins, out_channels, use_depth_distribution
        self.features = nn.Sequential(
            nn.Conv2d(3, 32, 3, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(32, 64, 3, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(64, width, 3, stride=2, padding=1), nn.ReLU(inplace=True))
        self.depth_layer = nn.Conv2d(width, out_channels + (depth_bins if use_depth_distribution else 0), kernel_size=1, padding=0)

    def get_features(self, x: torch.Tensor) -> torch.Tensor:
        return self.features(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.depth_layer(self.get_features(x))


class BevHead(nn.Module):
    """Stand-in consumer of the BEV features (temporal model + decoder of the reference are out of scope): segmentation (2 classes),
    centerness (1) and offset (2) maps at BEV resolution, the three outputs ``shared_step`` always trains (trainer.py:86-103)."""

    def __init__(self, in_channels: int, width: int = 32):
        super().__init__()
        self.trunk = nn.Sequential(nn.Conv2d(in_channels, width, 3, padding=1), nn.ReLU(inplace=True))
        self.segmentation = nn.Conv2d(width, 2, 1)
        self.centerness = nn.Conv2d(width, 1, 1)
        self.offset = nn.Conv2d(width, 2, 1)

    def forward(self, bev: torch.Tensor) -> Dict[str, torch.Tensor]:
        t = self.trunk(bev)
        return {"segmentation": self.segmentation(t), "instance_center": torch.sigmoid(self.centerness(t)), "instance_offset": self.offset(t)}


class LiftTrainModel(nn.Module):
    """images -> head tensor -> fused lift -> BEV head, with the reference's learnt loss weights (fiery.py:100-102 / trainer.py:86-103)."""

    def __init__(self, cfg: LiftConfig, feature_input: bool = False):
        super().__init__()
        self.cfg, self.feature_input = cfg, feature_input
        self.encoder = StandInEncoder(cfg.depth_bins, cfg.out_channels, cfg.use_depth_distribution)
        # the head tensor's producer on the tensor cores: half features in (AMP), fp32 head out -- no widening pass before the lift
        self.encoder.depth_layer = DepthLayer.from_conv(self.encoder.depth_layer)
        self.lift = LiftSplat.from_config(cfg, output_layout="channels_last")     # the BEV convs consume channels-last: no layout pass
        self.head = BevHead(cfg.out_channels)
        self.segmentation_weight = nn.Parameter(torch.tensor(0.0))
        self.centerness_weight = nn.Parameter(torch.tensor(0.0))
        self.offset_weight = nn.Parameter(torch.tensor(0.0))

    def forward(self, image: torch.Tensor, intrinsics: torch.Tensor, extrinsics: torch.Tensor) -> Dict[str, torch.Tensor]:
        """image (b, s, n, 3, H, W) [or, with ``feature_input``, the 128-channel feature maps (b, s, n, 128, h, w)], intrinsics
        (b, s, n, 3, 3), extrinsics (b, s, n, 4, 4) -> dict of (b, s, ., X, Y) maps: Fiery.calculate_birds_eye_view_features
        (fiery.py:275-286) followed by the head."""
        b, s, n = image.shape[:3]
        x = pack_sequence_dim(image)
        x = x.reshape(b * s * n, *x.shape[2:])
        head = self.encoder.depth_layer(x) if self.feature_input else self.encoder(x)      # encoder.py:94-96
        bev = self.lift(head, pack_sequence_dim(intrinsics), pack_sequence_dim(extrinsics))     # (b*s, C, X, Y) fp32
        out = self.head(bev)
        return {k: unpack_sequence_dim(v, b, s) for k, v in out.items()}

    def loss(self, output: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Uncertainty-weighted sum of trainer.py:86-103 (cross-entropy / L2 / L1 instead of the reference's top-k and masked variants)."""
        seg = output["segmentation"].float()
        b, s = seg.shape[:2]
        l_seg = F.cross_entropy(seg.reshape(b * s, *seg.shape[2:]), batch["segmentation"].reshape(b * s, *seg.shape[3:]))
        l_cen = F.mse_loss(output["instance_center"].float(), batch["centerness"])
        l_off = F.l1_loss(output["instance_offset"].float(), batch["offset"])
        return (l_seg / torch.exp(self.segmentation_weight) + 0.5 * self.segmentation_weight
                + l_cen / (2 * torch.exp(self.centerness_weight)) + 0.5 * self.centerness_weight
                + l_off / (2 * torch.exp(self.offset_weight)) + 0.5 * self.offset_weight)


def synthetic_batch(cfg: LiftConfig, batch: int, seq: int, device: torch.device, seed: int = 0, feature_input: bool = False,
                    first_sample: int = 0) -> Dict[str, torch.Tensor]:
    """A batch with the keys / shapes / dtypes of ``FuturePredictionDataset.__getitem__`` (fiery/data.py:343-363), batched:
    image (b, s, n, 3, H, W) float, intrinsics (b, s, n, 3, 3), extrinsics (b, s, n, 4, 4), segmentation (b, s, 1, X, Y) int64,
    centerness (b, s, 1, X, Y), offset (b, s, 2, X, Y), future_egomotion (b, s, 6).  Sample i of the global batch is generated from
    ``seed + first_sample + i`` only, so a rank's shard equals the same rows of the global batch."""
    H, W = cfg.final_dim
    h, w = cfg.feat_hw
    X, Y = cfg.bev_hw
    n = cfg.n_cameras
    out = {k: [] for k in ("image", "intrinsics", "extrinsics", "segmentation", "centerness", "offset", "future_egomotion")}
    for i in range(batch):
        sid = seed + first_sample + i
        rng = np.random.default_rng(sid + 7000003)
        one = LiftConfig(**{**cfg.__dict__, "frames": seq})
        K, E = make_calibration(one, seed=sid)
        shape = (seq, n, 128, h, w) if feature_input else (seq, n, 3, H, W)
        out["image"].append(torch.from_numpy(rng.standard_normal(shape, dtype=np.float32)))
        out["intrinsics"].append(torch.from_numpy(K))
        out["extrinsics"].append(torch.from_numpy(E))
        out["segmentation"].append(torch.from_numpy((rng.random((seq, 1, X, Y)) < 0.05).astype(np.int64)))
        out["centerness"].append(torch.from_numpy(rng.random((seq, 1, X, Y), dtype=np.float32)))
        out["offset"].append(torch.from_numpy(rng.standard_normal((seq, 2, X, Y), dtype=np.float32)))
        out["future_egomotion"].append(torch.from_numpy(make_egomotion(1, seq, seed=sid)[0]))
    return {k: torch.stack(v).to(device) for k, v in out.items()}


class LiftTrainer:
    """One data-parallel training step: forward under autocast (PRECISION 16 -> fp16 + loss scaling), backward through the fused
    lift, ONE all-reduce of the flat gradient, clip to GRAD_NORM_CLIP, Adam."""

    def __init__(self, cfg: LiftConfig, device: torch.device, precision: int = 16, feature_input: bool = False, seed: int = 0,
                 group=None):
        torch.manual_seed(seed)                                        # same initial weights on every rank (DDP broadcasts rank 0's)
        self.model = LiftTrainModel(cfg, feature_input=feature_input).to(device).to(memory_format=torch.channels_last)
        self.device, self.group, self.precision = device, group, precision
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=LR, weight_decay=WEIGHT_DECAY)     # trainer.py:254-260
        self.bucket = FlatGradBucket(self.model.parameters())
        self.scaler = torch.amp.GradScaler("cuda", enabled=(precision == 16))
        self.steps = 0

    def forward_backward(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Forward, backward and the gradient exchange: afterwards ``bucket.flat`` holds the (loss-scaled) gradient averaged over
        the data-parallel group."""
        self.bucket.zero()
        with torch.autocast("cuda", dtype=torch.float16, enabled=(self.precision == 16)):
            output = self.model(batch["image"], batch["intrinsics"], batch["extrinsics"])
            loss = self.model.loss(output, batch)
        self.scaler.scale(loss).backward()
        self.bucket.all_reduce_mean(self.group)                        # the step's only collective (scaled gradients: linear)
        return loss.detach()

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        loss = self.forward_backward(batch)
        self.scaler.unscale_(self.optimizer)
        torch.nn.utils.clip_grad_norm_(self.bucket.params, GRAD_NORM_CLIP)            # train.py:38 gradient_clip_val
        self.scaler.step(self.optimizer)
        self.scaler.update()
        self.steps += 1
        return loss


def rank_shard(global_batch: int, world_size: int, rank: int):
    """Samples of the global batch this rank trains on (DistributedSampler-style contiguous split)."""
    r = shard_frames(global_batch, world_size, rank)
    return r.start, len(r)
