"""ransac_flow_b200: B200-native (sm_100a) implementation of RANSAC-Flow's per-pair inference
hot path behind the reference's own Python API.

Modules mirror the reference's module names: ``outil`` (utils/outil.py), ``model``
(model/model.py), ``coarseAlignFeatMatch`` (the CoarseAlign variants), ``kornia_geometry``
(kornia.geometry.HomographyWarper), plus ``pipeline`` (PredFlowMask / pair loop / getFlow / KITTI two-level flow),
``results`` (the drivers' .npy formats and the getResults metrics),
``shard`` (pair sharding over GPUs) and ``dropin`` (run the reference's scripts unchanged).
All arithmetic is in ``libransacflow_b200.so`` (C ABI: include/ransacflow_b200.h).
Importing fails if that library is missing; compute calls fail without a CUDA device.
"""
from . import _lib  # noqa: F401  (loads the C-ABI library or raises)
from . import ops, outil, model, kornia_geometry, coarseAlignFeatMatch, pipeline, results  # noqa: F401
from .coarseAlignFeatMatch import CoarseAlign, CoarseAlignA, CoarseAlignB, CoarseAlignC  # noqa: F401

__version__ = "0.1.0"
