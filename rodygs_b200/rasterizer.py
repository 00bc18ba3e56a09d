"""Drop-in replacement for the two names RoDyGS imports from `diff_gauss_pose`
(/root/reference/src/trainer/renderer.py:14, src/model/rodygs_static.py:19,
src/evaluator/eval.py:25):

    GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, bg,
                                  scale_modifier, projmatrix, sh_degree, prefiltered,
                                  debug, enable_cov_grad, enable_sh_grad)   # renderer.py:50-63
    GaussianRasterizer(raster_settings=...)(means3D=, means2D=, shs=, colors_precomp=,
                                  opacities=, scales=, rotations=, cov3Ds_precomp=,
                                  viewmatrix=)                               # renderer.py:65-101
    -> (color[3,H,W], depth[1,H,W], normal[3,H,W], alpha[1,H,W], radii[N] int32, extra)

Same names, argument meaning and error behaviour; autograd reaches means3D,
means2D (gradient sink, per NDC unit), shs / colors_precomp, opacities, scales,
rotations and viewmatrix.  `normal` is returned as zeros and `extra` as None:
no consumer exists in the reference.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import engine
from .engine import SceneArgs, SceneGrads, SetArgs, SetGrads, ViewArgs


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    projmatrix: torch.Tensor
    sh_degree: int
    prefiltered: bool = False
    debug: bool = False
    enable_cov_grad: bool = True
    enable_sh_grad: bool = True


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, viewmatrix,
                settings: GaussianRasterizationSettings):
        means3D, shs, colors_precomp = _f32c(means3D), _f32c(shs), _f32c(colors_precomp)
        opacities, scales, rotations = _f32c(opacities), _f32c(scales), _f32c(rotations)
        viewmatrix = _f32c(viewmatrix)
        n = means3D.shape[0]
        sh_coeffs = 0 if shs is None else shs.shape[1]
        if shs is not None and (shs.dim() != 3 or shs.shape[2] != 3 or sh_coeffs not in (1, 16)
                                or sh_coeffs < (settings.sh_degree + 1) ** 2):
            raise Exception("shs must be [N,16,3] (get_features, rodygs_static.py:98-101) or [N,1,3] at sh_degree 0")
        st = SetArgs(xyz=means3D, scaling=scales, rotation=rotations, opacity=opacities,
                     sh_dc=shs, sh_rest=shs, sh_dc_stride=3 * sh_coeffs, sh_rest_stride=3 * sh_coeffs,
                     sh_rest_offset=3)
        scene = SceneArgs(st=st, raw=False, colors_precomp=colors_precomp)
        view = ViewArgs(height=settings.image_height, width=settings.image_width, tanfovx=settings.tanfovx,
                        tanfovy=settings.tanfovy, scale_modifier=settings.scale_modifier,
                        sh_degree=settings.sh_degree, viewmatrix=viewmatrix,
                        projmatrix=_f32c(settings.projmatrix), bg=_f32c(settings.bg),
                        enable_cov_grad=bool(settings.enable_cov_grad), enable_sh_grad=bool(settings.enable_sh_grad))
        color, depth, alpha, radii, state = engine.render_forward(scene, view)
        ctx.state = state
        ctx.n = n
        ctx.sh_coeffs = sh_coeffs
        ctx.mark_non_differentiable(radii)
        return color, depth, alpha, radii

    @staticmethod
    def backward(ctx, dL_dcolor, dL_ddepth, dL_dalpha, _dradii):
        state = ctx.state
        n = ctx.n
        dev = state.view.viewmatrix.device
        f32 = dict(dtype=torch.float32, device=dev)
        scene = state.scene
        g_shs = torch.empty(n, ctx.sh_coeffs, 3, **f32) if scene.st.sh_dc is not None else None
        if g_shs is not None and ctx.sh_coeffs > 16:
            g_shs.zero_()
        grads = SceneGrads(
            st=SetGrads(xyz=torch.empty(n, 3, **f32), scaling=torch.empty(n, 3, **f32),
                        rotation=torch.empty(n, 4, **f32), opacity=torch.empty(n, 1, **f32),
                        sh_dc=g_shs, sh_rest=g_shs if (g_shs is not None and ctx.sh_coeffs > 1) else None,
                        sh_rest_offset=3),
            colors_precomp=torch.empty(n, 3, **f32) if scene.colors_precomp is not None else None,
            means2D=torch.empty(n, 3, **f32),
            viewmatrix=torch.zeros(4, 4, **f32),
        )
        engine.render_backward(state, dL_dcolor, dL_ddepth, dL_dalpha, grads)
        ctx.state = None
        return (grads.st.xyz, grads.means2D, g_shs, grads.colors_precomp, grads.st.opacity, grads.st.scaling,
                grads.st.rotation, grads.viewmatrix, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3Ds_precomp=None, viewmatrix=None, cov3D_precomp=None):
        if cov3Ds_precomp is None:
            cov3Ds_precomp = cov3D_precomp
        # same checks and messages as the upstream Python wrapper
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3Ds_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3Ds_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if cov3Ds_precomp is not None:
            raise NotImplementedError(
                "cov3Ds_precomp is not supported: RoDyGS always passes scales + rotations (renderer.py:69-74)")
        if viewmatrix is None:
            raise Exception("viewmatrix (world_view_transform, transposed) is required (renderer.py:97-99)")
        st = self.raster_settings
        color, depth, alpha, radii = _RasterizeGaussians.apply(
            means3D, means2D, shs, colors_precomp, opacities, scales, rotations, viewmatrix, st)
        normal = torch.zeros(3, int(st.image_height), int(st.image_width), dtype=color.dtype, device=color.device)
        return color, depth, normal, alpha, radii, None
