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

import importlib

from .geometry import VoxelsSumming
from .lift import calculate_birds_eye_view_features
from .warp import cumulative_warp_features, warp_features

_saved = {}


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
        fiery_mod.Fiery.calculate_birds_eye_view_features = calculate_birds_eye_view_features
    if level == "all":
        geometry.cumulative_warp_features = cumulative_warp_features
        geometry.warp_features = warp_features
        fiery_mod.cumulative_warp_features = cumulative_warp_features
    return fiery_mod.Fiery


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
