"""ctypes binding of libfiery_b200.so (C ABI: include/fiery_b200.h).

There is no fallback: if the shared library is missing or a call fails this module raises.  Build it in-tree with
``python -m fiery_b200.build`` (the built ``.so`` travels with the repo snapshot to the GPU box).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint8, c_void_p

# FIERY_B200_LIB: another build of the same library (experiment builds of tools/gpu_ab.sh); default: the in-tree one
LIB_PATH = os.environ.get("FIERY_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfiery_b200.so")
ABI_VERSION = 2

DTYPE_F32, DTYPE_F16 = 0, 1
CALIB_RAW, CALIB_COMPOSED = 0, 1
BEV_NCHW, BEV_NHWC = 0, 1


class FieryError(RuntimeError):
    """A fiery_b200 C-ABI call returned a negative status."""


class LiftDesc(ctypes.Structure):
    """Mirror of ``fiery_lift_desc_t``."""

    _fields_ = [
        ("n_frames", c_int32), ("n_cameras", c_int32), ("depth_bins", c_int32), ("channels", c_int32),
        ("feat_h", c_int32), ("feat_w", c_int32),
        ("bev_x", c_int32), ("bev_y", c_int32), ("bev_z", c_int32),
        ("bev_offset", c_float * 3), ("bev_resolution", c_float * 3),
        ("z_valid_lo", c_float), ("z_valid_hi", c_float),
        ("use_depth_distribution", c_int32), ("head_dtype", c_int32), ("calib_mode", c_int32), ("bev_layout", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/fiery_b200.h declares
SIGNATURES = {
    "fiery_abi_version": (c_int32, []),
    "fiery_last_error": (c_char_p, []),
    "fiery_lift_plan_bytes": (c_size_t, [POINTER(LiftDesc)]),
    "fiery_lift_plan": (c_int32, [POINTER(LiftDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fiery_lift_scratch_bytes": (c_size_t, [POINTER(LiftDesc)]),
    "fiery_lift_forward_launches": (c_int32, [POINTER(LiftDesc)]),
    "fiery_lift_forward": (c_int32, [POINTER(LiftDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "fiery_lift_forward_warped": (c_int32, [POINTER(LiftDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fiery_lift_forward_timed": (c_int32, [POINTER(LiftDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_int32, POINTER(c_float), POINTER(c_int32),
                                           POINTER(c_int32)]),
    "fiery_lift_set_max_chunk_frames": (None, [c_int32]),
    "fiery_lift_workspace_bytes": (c_size_t, [POINTER(LiftDesc)]),
    "fiery_lift_backward": (c_int32, [POINTER(LiftDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fiery_lift_point_indices": (c_int32, [POINTER(LiftDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "fiery_compose_calibration": (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fiery_voxels_summing_plan": (c_int32, [c_int64, c_void_p, c_void_p, POINTER(c_int64), c_void_p]),
    "fiery_voxels_summing_forward": (c_int32, [c_int64, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                               c_void_p, c_void_p, c_void_p]),
    "fiery_voxels_summing_backward": (c_int32, [c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fiery_depth_layer_forward": (c_int32, [c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fiery_bev_conv_pack_weights": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "fiery_bev_first_conv_forward": (c_int32, [c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "fiery_warp_theta": (c_int32, [c_int32, c_int32, c_int32, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "fiery_warp_features_forward": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                              c_int64, c_int32, c_void_p]),
    "fiery_warp_features_backward": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                               c_int64, c_int32, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Loads the shared library once; raises ``FieryError`` if it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FieryError(f"{LIB_PATH} not found: build it with `python -m fiery_b200.build` "
                         "(fiery_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.fiery_abi_version()
    if got != ABI_VERSION:
        raise FieryError(f"libfiery_b200.so has ABI version {got}, this package expects {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().fiery_last_error()
        raise FieryError(f"{what} failed ({status}): {msg.decode() if msg else 'no message'}")
