"""Import shim: `from diff_gauss_pose import GaussianRasterizationSettings, GaussianRasterizer`
(/root/reference/src/trainer/renderer.py:14, src/model/rodygs_static.py:19,
src/evaluator/eval.py:25) resolves to the B200-native implementation."""
from rodygs_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
