"""Rebind the reference's call sites to the CUDA path.

``install()`` patches an importable ``fiery`` package (the unmodified reference) so that ``Fiery.forward``
(fiery/models/fiery.py:130-191) runs on the fused lift without any source change:

  * ``fiery.utils.geometry.VoxelsSumming`` and ``fiery.models.fiery.VoxelsSumming`` (the name is bound at import,
    fiery.py:10) -> ``fiery_b200.geometry.VoxelsSumming``                                   [level "voxels_summing"]
  * ``Fiery.calculate_birds_eye_view_features`` (fiery.py:275) -> ``fiery_b200.lift.calculate_birds_eye_view_features``
                                                                                             [level "fused", default]
"""
from __future__ import annotations

import importlib

from .geometry import VoxelsSumming
from .lift import calculate_birds_eye_view_features

_saved = {}


def install(level: str = "fused"):
    if level not in ("fused", "voxels_summing"):
        raise ValueError("level must be 'fused' or 'voxels_summing'")
    geometry = importlib.import_module("fiery.utils.geometry")
    fiery_mod = importlib.import_module("fiery.models.fiery")
    if not _saved:
        _saved["VoxelsSumming"] = geometry.VoxelsSumming
        _saved["bev"] = fiery_mod.Fiery.calculate_birds_eye_view_features
    geometry.VoxelsSumming = VoxelsSumming
    fiery_mod.VoxelsSumming = VoxelsSumming
    if level == "fused":
        fiery_mod.Fiery.calculate_birds_eye_view_features = calculate_birds_eye_view_features
    return fiery_mod.Fiery


def uninstall():
    if not _saved:
        return
    geometry = importlib.import_module("fiery.utils.geometry")
    fiery_mod = importlib.import_module("fiery.models.fiery")
    geometry.VoxelsSumming = _saved["VoxelsSumming"]
    fiery_mod.VoxelsSumming = _saved["VoxelsSumming"]
    fiery_mod.Fiery.calculate_birds_eye_view_features = _saved["bev"]
    _saved.clear()
