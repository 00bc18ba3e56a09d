"""Fused time-conditioned render on the *raw* parameters of the static and the
dynamic model - the B200-first replacement of the reference's per-step chain

    DynRoDyGS.get_gaussian_deformation(t)      /root/reference/src/model/rodygs_dynamic.py:122-138
    activation getters                         /root/reference/src/model/rodygs_static.py:82-105
    RoDyGSTrainer.get_GS_properties (5x cat)   /root/reference/src/trainer/rodygs.py:68-113
    render(...)                                /root/reference/src/trainer/renderer.py:17-114

Nothing is concatenated or materialised: the two models are two base pointers,
exp / normalize / sigmoid and  c_i . (B(t) - B(t_i))  happen in registers inside the
preprocess kernel, and the backward kernel writes gradients of the raw parameters,
of the motion coefficients, of B(t) and of the motion table directly.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from . import engine
from .engine import SceneArgs, SceneGrads, SetArgs, SetGrads, ViewArgs
from .rasterizer import GaussianRasterizationSettings, _f32c


class GaussianParams(NamedTuple):
    """Un-activated parameters of one model, named as in rodygs_static.py:35-47."""
    xyz: torch.Tensor            # [n,3]
    features_dc: torch.Tensor    # [n,1,3]
    features_rest: torch.Tensor  # [n,15,3]
    scaling: torch.Tensor        # [n,3]  log-scale
    rotation: torch.Tensor       # [n,4]  raw quaternion (r,x,y,z)
    opacity: torch.Tensor        # [n,1]  logit


def _set_of(p: Optional[GaussianParams]) -> Optional[SetArgs]:
    if p is None or p.xyz.shape[0] == 0:
        return None
    if p.features_rest.shape[1:] != (15, 3) or p.features_dc.shape[1:] != (1, 3):
        raise Exception("features_dc must be [n,1,3] and features_rest [n,15,3] (max_sh_degree 3)")
    if p.scaling.shape[1] != 3:
        raise NotImplementedError("isotropic models (scaling [n,1]) are not supported by the fused path")
    return SetArgs(xyz=_f32c(p.xyz), scaling=_f32c(p.scaling), rotation=_f32c(p.rotation), opacity=_f32c(p.opacity),
                   sh_dc=_f32c(p.features_dc), sh_rest=_f32c(p.features_rest), sh_dc_stride=3, sh_rest_stride=45,
                   sh_rest_offset=0)


def _grads_like(p: Optional[SetArgs], dev) -> SetGrads:
    if p is None:
        return SetGrads()
    n = p.n()
    f32 = dict(dtype=torch.float32, device=dev)
    return SetGrads(xyz=torch.empty(n, 3, **f32), scaling=torch.empty(n, 3, **f32), rotation=torch.empty(n, 4, **f32),
                    opacity=torch.empty(n, 1, **f32), sh_dc=torch.empty(n, 1, 3, **f32),
                    sh_rest=torch.empty(n, 15, 3, **f32))


class _FusedDynamicRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, means2D, viewmatrix, motion_coeff, basis_t, table,
                s_xyz, s_fdc, s_frest, s_scal, s_rot, s_op,
                d_xyz, d_fdc, d_frest, d_scal, d_rot, d_op):
        settings, time_ind, spatial_lr_scale, use_deform = meta
        st = _set_of(GaussianParams(s_xyz, s_fdc, s_frest, s_scal, s_rot, s_op)) if s_xyz is not None else None
        dy = _set_of(GaussianParams(d_xyz, d_fdc, d_frest, d_scal, d_rot, d_op)) if d_xyz is not None else None
        nd = dy.n() if dy is not None else 0
        deform = bool(use_deform) and nd > 0
        coeff = None
        if deform:
            coeff = _f32c(motion_coeff).reshape(nd, -1)
            basis_t, table = _f32c(basis_t), _f32c(table)
            if coeff.shape[1] > 16 or basis_t.shape != (coeff.shape[1], 7) or table.shape[1:] != (coeff.shape[1], 7):
                raise Exception("motion_coeff [Nd,(1,)K<=16], basis_t [K,7] and table [T,K,7] are required")
        order = offsets = None
        if deform:
            time_ind, order, offsets = engine.frame_csr_of(time_ind, table.shape[0])
        scene = SceneArgs(st=st, dy=dy, raw=True, use_deform=deform, motion_coeff=coeff, frame_order=order, frame_offsets=offsets,
                          time_ind=time_ind if deform else None, basis_t=basis_t if deform else None,
                          table=table if deform else None, spatial_lr_scale=float(spatial_lr_scale))
        view = ViewArgs(height=settings.image_height, width=settings.image_width, tanfovx=settings.tanfovx,
                        tanfovy=settings.tanfovy, scale_modifier=settings.scale_modifier, sh_degree=settings.sh_degree,
                        viewmatrix=_f32c(viewmatrix), projmatrix=_f32c(settings.projmatrix), bg=_f32c(settings.bg),
                        enable_cov_grad=bool(settings.enable_cov_grad), enable_sh_grad=bool(settings.enable_sh_grad))
        color, depth, alpha, radii, state = engine.render_forward(scene, view)
        ctx.state = state
        ctx.coeff_shape = None if motion_coeff is None else motion_coeff.shape
        ctx.has = (s_xyz is not None, d_xyz is not None)
        ctx.mark_non_differentiable(radii)
        return color, depth, alpha, radii

    @staticmethod
    def backward(ctx, dL_dcolor, dL_ddepth, dL_dalpha, _dradii):
        state = ctx.state
        scene = state.scene
        dev = state.view.viewmatrix.device
        f32 = dict(dtype=torch.float32, device=dev)
        grads = SceneGrads(st=_grads_like(scene.st, dev), dy=_grads_like(scene.dy, dev),
                           means2D=torch.empty(state.n, 3, **f32), viewmatrix=torch.zeros(4, 4, **f32))
        if scene.use_deform:
            grads.motion_coeff = torch.empty_like(scene.motion_coeff)
            grads.table = torch.zeros_like(scene.table)
            grads.basis_t = torch.zeros_like(scene.basis_t)
        engine.render_backward(state, dL_dcolor, dL_ddepth, dL_dalpha, grads)
        ctx.state = None
        gc = None if grads.motion_coeff is None else grads.motion_coeff.reshape(ctx.coeff_shape)
        s, d = grads.st, grads.dy
        return (None, grads.means2D, grads.viewmatrix, gc, grads.basis_t, grads.table,
                s.xyz, s.sh_dc, s.sh_rest, s.scaling, s.rotation, s.opacity,
                d.xyz, d.sh_dc, d.sh_rest, d.scaling, d.rotation, d.opacity)


def render_dynamic(static: Optional[GaussianParams], dynamic: Optional[GaussianParams],
                   settings: GaussianRasterizationSettings, viewmatrix: torch.Tensor,
                   motion_coeff: Optional[torch.Tensor] = None, basis_t: Optional[torch.Tensor] = None,
                   table: Optional[torch.Tensor] = None, time_ind: Optional[torch.Tensor] = None,
                   spatial_lr_scale: float = 1.0, use_deform: bool = True, means2D: Optional[torch.Tensor] = None):
    """One fused render.  Returns the dict of renderer.py:103-114 (same keys).

    basis_t = B(t) [K,7] and table = B at every training time [T,K,7] are the outputs
    of the motion-basis MLP (rodygs_b200.deform.MotionBasisNetwork, still PyTorch);
    time_ind [Nd] is `gaussian_to_time_ind` (rodygs_dynamic.py:58-77)."""
    ns = 0 if static is None else static.xyz.shape[0]
    nd = 0 if dynamic is None else dynamic.xyz.shape[0]
    ref = static.xyz if ns > 0 else dynamic.xyz
    if means2D is None:
        means2D = torch.zeros(ns + nd, 3, dtype=torch.float32, device=ref.device, requires_grad=True)
    s = tuple(static) if ns > 0 else (None,) * 6
    d = tuple(dynamic) if nd > 0 else (None,) * 6
    meta = (settings, time_ind, spatial_lr_scale, use_deform)
    color, depth, alpha, radii = _FusedDynamicRender.apply(meta, means2D, viewmatrix, motion_coeff, basis_t, table, *s, *d)
    H, W = int(settings.image_height), int(settings.image_width)
    return {
        "rendered_image": color,
        "rendered_depth": depth,
        "rendered_normal": torch.zeros(3, H, W, dtype=color.dtype, device=color.device),
        "rendered_alpha": alpha,
        "viewspace_points": means2D,
        "visibility_filter": radii > 0,
        "radii": radii,
        "extra": None,
    }
