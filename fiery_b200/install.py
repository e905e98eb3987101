"""Rebind the reference's call sites to the CUDA path.

``install()`` patches an importable ``fiery`` package (the unmodified reference) so that ``Fiery.forward``
(fiery/models/fiery.py:130-191) runs on the fused lift without any source change:

  * ``fiery.utils.geometry.VoxelsSumming`` and ``fiery.models.fiery.VoxelsSumming`` (the name is bound at import,
    fiery.py:10) -> ``fiery_b200.geometry.VoxelsSumming``                                   [level "voxels_summing"]
  * ``Fiery.calculate_birds_eye_view_features`` (fiery.py:275) -> ``fiery_b200.lift.calculate_birds_eye_view_features``
                                                                                             [level "fused", default]
  * ``fiery.models.fiery.cumulative_warp_features`` (bound at import, fiery.py:10; call site fiery.py:143) and
    ``fiery.utils.geometry.cumulative_warp_features`` / ``warp_features`` -> ``fiery_b200.warp``       [level "all"]
"""
from __future__ import annotations

import functools
import importlib
import warnings

from .geometry import VoxelsSumming
from .lift import calculate_birds_eye_view_features
from .warp import cumulative_warp_features, warp_features

_saved = {}
_warned = set()


def unsupported_reason(model, x):
    """None if the fused kernels cover this model's lift configuration, else the reason (the limits of
    fiery_b200/csrc: include/fiery_b200.h).  CPU tensors are NOT a reason: there is no CPU path and the call raises."""
    h, w = x.shape[-2] // model.encoder_downsample, x.shape[-1] // model.encoder_downsample
    D = model.frustum.shape[0]
    if int(model.encoder_out_channels) != 64:
        return f"MODEL.ENCODER.OUT_CHANNELS={int(model.encoder_out_channels)} (kernels are built for 64)"
    if D > 48:
        return f"{D} depth bins (kernels are built for <= 48)"
    if h > 32 or w % 4:
        return f"feature map {h}x{w} (kernels need h <= 32 and w % 4 == 0)"
    if int(model.bev_dimension[2]) != 1:
        return "more than one height cell"
    return None


def _warn_once(key, msg):
    if key not in _warned:
        _warned.add(key)
        warnings.warn(msg, RuntimeWarning, stacklevel=3)


def _bev_features(self, x, intrinsics, extrinsics):
    """Installed as ``Fiery.calculate_birds_eye_view_features``: the fused lift where the kernels cover the configuration,
    the reference's own method (unpatched behaviour, its own device) where they do not."""
    reason = unsupported_reason(self, x)
    if reason is None:
        return calculate_birds_eye_view_features(self, x, intrinsics, extrinsics)
    _warn_once(("bev", reason), f"fiery_b200: lift configuration not covered by the CUDA kernels ({reason}); "
                                "running the reference's own calculate_birds_eye_view_features")
    return _saved["bev"](self, x, intrinsics, extrinsics)


def _bilinear_only(ours, saved_key):
    """Feature warps (mode='bilinear', fiery.py:143) run on the CUDA kernel.  The trainer's label warps use mode='nearest'
    (trainer.py, cumulative_warp_features_reverse): sample positions agree with torch only to ~1e-4, which can flip a
    nearest-neighbour pick at a tie, so those stay on the reference's function."""
    @functools.wraps(ours)
    def fn(x, flow, mode="nearest", spatial_extent=None):
        if mode == "bilinear" or _saved.get(saved_key) is None:
            return ours(x, flow, mode=mode, spatial_extent=spatial_extent)
        return _saved[saved_key](x, flow, mode=mode, spatial_extent=spatial_extent)
    return fn


def install(level: str = "fused"):
    if level not in ("fused", "voxels_summing", "all"):
        raise ValueError("level must be 'fused', 'voxels_summing' or 'all'")
    geometry = importlib.import_module("fiery.utils.geometry")
    fiery_mod = importlib.import_module("fiery.models.fiery")
    if not _saved:
        _saved["VoxelsSumming"] = geometry.VoxelsSumming
        _saved["bev"] = fiery_mod.Fiery.calculate_birds_eye_view_features
        _saved["cwf"] = getattr(geometry, "cumulative_warp_features", None)
        _saved["wf"] = getattr(geometry, "warp_features", None)
        _saved["cwf_model"] = getattr(fiery_mod, "cumulative_warp_features", None)
    geometry.VoxelsSumming = VoxelsSumming
    fiery_mod.VoxelsSumming = VoxelsSumming
    if level in ("fused", "all"):
        fiery_mod.Fiery.calculate_birds_eye_view_features = _bev_features
    if level == "all":
        geometry.cumulative_warp_features = _bilinear_only(cumulative_warp_features, "cwf")
        geometry.warp_features = _bilinear_only(warp_features, "wf")
        fiery_mod.cumulative_warp_features = _bilinear_only(cumulative_warp_features, "cwf_model")
    return fiery_mod.Fiery


def use_tensor_core_depth_layer(model):
    """Replace ``model.encoder.depth_layer`` (``nn.Conv2d(128, D + C, 1)``, fiery/models/encoder.py:36) of a ``Fiery`` instance by
    ``fiery_b200.depth_layer.DepthLayer`` sharing the same Parameters (``state_dict`` keys unchanged): under autocast the backbone's
    half features go straight to an fp32 head tensor, the dtype the lift computes in.  Returns the model; layers the kernel does not
    cover (input channels != 128, more than 128 outputs) are left alone with one warning."""
    from .depth_layer import DepthLayer
    conv = model.encoder.depth_layer
    if isinstance(conv, DepthLayer):
        return model
    if conv.in_channels != 128 or conv.out_channels > 128 or conv.kernel_size != (1, 1):
        _warn_once(("depth_layer", conv.in_channels, conv.out_channels),
                   f"fiery_b200: depth_layer {conv.in_channels}->{conv.out_channels} not covered by the tensor-core kernel; left as is")
        return model
    model.encoder.depth_layer = DepthLayer.from_conv(conv)
    return model


def uninstall():
    if not _saved:
        return
    geometry = importlib.import_module("fiery.utils.geometry")
    fiery_mod = importlib.import_module("fiery.models.fiery")
    geometry.VoxelsSumming = _saved["VoxelsSumming"]
    fiery_mod.VoxelsSumming = _saved["VoxelsSumming"]
    fiery_mod.Fiery.calculate_birds_eye_view_features = _saved["bev"]
    for mod, name, key in ((geometry, "cumulative_warp_features", "cwf"), (geometry, "warp_features", "wf"),
                           (fiery_mod, "cumulative_warp_features", "cwf_model")):
        if _saved.get(key) is not None:
            setattr(mod, name, _saved[key])
    _saved.clear()
    _warned.clear()
