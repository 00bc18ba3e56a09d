"""rodygs_b200 - B200-native (sm_100a) implementation of RoDyGS's per-iteration
dynamic-splatting hot path behind the `diff_gauss_pose` API.

Importing this package does not need a GPU; calling into it does, and needs the
in-tree CUDA library (`python -m rodygs_b200.build`).  There is no CPU fallback.
"""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401
from .engine import config  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "config"]
