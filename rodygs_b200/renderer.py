"""`render(...)` with the reference's signature and return dict
(/root/reference/src/trainer/renderer.py:17-114), running on the B200-native
rasterizer.  A maintainer can point `src.trainer.rodygs` at this function, or
simply make `diff_gauss_pose` resolve to the shim at the repo root."""
from __future__ import annotations

import math

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def render(xyz, active_sh_degree, opacity, scaling, rotation, features, viewpoint_camera, bg_color: torch.Tensor,
           scaling_modifier=1, override_color=None, enable_sh_grad=False, enable_cov_grad=False):
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    # the reference passes enable_cov_grad=enable_sh_grad and vice versa (renderer.py:61-62);
    # both are always equal at its call sites (rodygs.py:268-269), the swap is kept for fidelity.
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        projmatrix=viewpoint_camera.projection_matrix.transpose(0, 1), sh_degree=active_sh_degree,
        prefiltered=False, debug=False, enable_cov_grad=enable_sh_grad, enable_sh_grad=enable_cov_grad)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    shs, colors_precomp = (features, None) if override_color is None else (None, override_color)
    image, depth, normal, alpha, radii, extra = rasterizer(
        means3D=xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp, opacities=opacity,
        scales=scaling, rotations=rotation, cov3Ds_precomp=None,
        viewmatrix=viewpoint_camera.world_view_transform.transpose(0, 1))
    return {"rendered_image": image, "rendered_depth": depth, "rendered_normal": normal, "rendered_alpha": alpha,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii, "extra": extra}
