"""PyTorch-CPU restatement of the tile rasterizer behind
``GaussianRasterizationSettings`` / ``GaussianRasterizer``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**:
the reference's rasterizer source is not in ``/root/reference`` (un-vendored
submodule ``diff_gauss_pose``), so this file follows SURVEY.md Appendix A
(public 3DGS algorithm + depth/alpha outputs + pose gradient), anchored on
the call sites ``/root/reference/src/trainer/renderer.py:50-101`` (argument
meaning, "glm storage" transposes, 6-tuple return) and
``/root/reference/src/utils/sh_utils.py:24-41,72-101`` (SH constants/signs).

Arithmetic contract (what "bit-exact" in the tests means)
---------------------------------------------------------
Every quantity that feeds an integer output (depth bits -> sort keys, pixel
centre / radius -> tile rectangle -> tile ranges, radii) is computed in
float32 with *individually rounded* IEEE operations in a fixed left-to-right
order; PyTorch's element-wise CPU kernels give exactly that (one op, one
rounding, no FMA contraction, no reassociation).  The CUDA side compiles the
same expressions with ``-fmad=false``.  No matmul / sum reductions are used on
that path (their summation order is unspecified).  ``exp`` only feeds the
1e-4-tolerance outputs.

Backward is PyTorch autograd over this forward, with the upstream
conventions of SURVEY.md App. A.6 written in explicitly:
(i) straight-through ``min(0.99, .)``; (ii) skip rules are constants;
(iii) SH clamp zeroes the gradient; (iv) the +-1.3 tan(fov) clamp freezes the
clamped coordinate; (v) the cov2D-inverse backward divides by ``det^2 + 1e-7``
(``_ConicFromCov``); (vi) no quaternion-normalisation Jacobian.

The outputs of this file on a fixed scene are frozen in
``tests/golden/raster_small.npz`` (``tests/golden/make_golden_raster.py``), and
``oracle/splat_scalar.py`` is a second, independently structured restatement
(scalar loops) that ``tests/test_oracle_scalar.py`` checks this one against.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch

TILE = 16
NEAR_Z = 0.2
LOWPASS = 0.3
ALPHA_MAX = 0.99
ALPHA_MIN = 1.0 / 255.0
T_STOP = 1e-4

# /root/reference/src/utils/sh_utils.py:24-41
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
         0.3731763325901154, -0.4570457994644658, 1.445305721320277,
         -0.5900435899266435]


class Settings(NamedTuple):
    """Mirror of the 12 keyword fields at renderer.py:50-63."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    projmatrix: torch.Tensor  # P^T ("glm storage")
    sh_degree: int
    prefiltered: bool = False
    debug: bool = False
    enable_cov_grad: bool = True
    enable_sh_grad: bool = True


def sh_to_rgb(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """sh: [M, K, 3] (coefficient-major, like ``get_features``), dirs: [M,3]
    unit vectors.  Returns the *unclamped* SH sum [M,3] (no +0.5).
    Basis and signs: /root/reference/src/utils/sh_utils.py:72-101."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            res = (res
                   + SH_C2[0] * xy * sh[:, 4]
                   + SH_C2[1] * yz * sh[:, 5]
                   + SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6]
                   + SH_C2[3] * xz * sh[:, 7]
                   + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res
                       + SH_C3[0] * y * (3.0 * xx - yy) * sh[:, 9]
                       + SH_C3[1] * xy * z * sh[:, 10]
                       + SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[:, 11]
                       + SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12]
                       + SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[:, 13]
                       + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                       + SH_C3[6] * x * (xx - 3.0 * yy) * sh[:, 15])
    return res


def sh_basis(deg: int, dirs: torch.Tensor) -> torch.Tensor:
    """Y_k(dir) for k < (deg+1)^2 as [M,K]: the coefficients sh_to_rgb multiplies (sh_utils.py:72-101)."""
    K = (deg + 1) ** 2
    eye = torch.eye(K, dtype=dirs.dtype)
    cols = [sh_to_rgb(deg, eye[k].reshape(1, K, 1).expand(dirs.shape[0], K, 3), dirs)[:, 0] for k in range(K)]
    return torch.stack(cols, 1)


def sh_grad_from_factors(deg: int, means_per_view, campos_per_view, dcolor_per_view) -> torch.Tensor:
    """dL/dSH [N,16,3] of several views from its rank-1 factors (the data-parallel exchange of
    rodygs_b200/csrc/sh_grad_views.cu): sum_v Y_k(normalize(x_v - campos_v)) * dcolor_v[c], where dcolor_v is
    dL/d(rgb) after the clamp mask (zero for Gaussians that view v does not see)."""
    n = means_per_view[0].shape[0]
    K = (deg + 1) ** 2
    g = torch.zeros(n, 16, 3, dtype=means_per_view[0].dtype)
    for x, cp, dc in zip(means_per_view, campos_per_view, dcolor_per_view):
        d = x - cp.reshape(1, 3)
        d = d / d.norm(dim=1, keepdim=True)
        g[:, :K] += sh_basis(deg, d).unsqueeze(2) * dc.unsqueeze(1)
    return g


class _ConicFromCov(torch.autograd.Function):
    """conic = cov2D^-1 = (c, -b, a) / det.  Forward exactly as App. A.2 step 5; backward as App. A.6 (v): the public
    implementation differentiates the inverse with ``1 / (det^2 + 1e-7)`` in place of ``1 / det^2`` (the B gradient is
    that of the single stored off-diagonal, which appears once in ``power``)."""

    @staticmethod
    def forward(ctx, ca, cb, cc, det, det_inv):
        ctx.save_for_backward(ca, cb, cc, det)
        return cc * det_inv, -cb * det_inv, ca * det_inv

    @staticmethod
    def backward(ctx, gA, gB, gC):
        ca, cb, cc, det = ctx.saved_tensors
        d2 = 1.0 / (det * det + 1e-7)
        da = d2 * (-cc * cc * gA + cb * cc * gB - cb * cb * gC)
        db = d2 * (2.0 * cb * cc * gA - (det + 2.0 * cb * cb) * gB + 2.0 * ca * cb * gC)
        dc = d2 * (-cb * cb * gA + ca * cb * gB - ca * ca * gC)
        return da, db, dc, None, None


class Preprocessed(NamedTuple):
    idx: torch.Tensor          # [M] global ids of Gaussians in front of the near plane
    visible: torch.Tensor      # [M] bool (survived det / empty-rect culls)
    radii: torch.Tensor        # [N] int32
    tiles_touched: torch.Tensor  # [N] int32
    rect_min: torch.Tensor     # [M,2] int32 (x,y)
    rect_max: torch.Tensor     # [M,2] int32
    xy: torch.Tensor           # [M,2] pixel coords (differentiable)
    depth: torch.Tensor        # [M]   view-space z (differentiable)
    conic: torch.Tensor        # [M,3] (A,B,C)
    opacity: torch.Tensor      # [M]
    rgb: torch.Tensor          # [M,3]
    cov2d: torch.Tensor        # [M,3] (a,b,c) after low-pass, for diagnostics


def _mat(m_t: torch.Tensor):
    """Tensor passed in 'glm storage' (transposed): element [i][j] = M[j][i]."""
    return m_t.t()


def preprocess(means3D, scales, rotations, opacities, shs, colors_precomp,
               viewmatrix, st: Settings, means2D=None) -> Preprocessed:
    """SURVEY.md App. A.2, steps 1-10.  All per-Gaussian, element-wise."""
    dt = means3D.dtype
    N = means3D.shape[0]
    H, W = int(st.image_height), int(st.image_width)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    V = _mat(viewmatrix.to(dt))
    P = _mat(st.projmatrix.to(dt))
    tanx = torch.tensor(st.tanfovx, dtype=dt)
    tany = torch.tensor(st.tanfovy, dtype=dt)

    def c(v):
        return torch.tensor(v, dtype=dt)

    # -- 1. view-space position, near-plane cull --------------------------------
    x, y, z = means3D[:, 0], means3D[:, 1], means3D[:, 2]
    tz_all = V[2, 0] * x + V[2, 1] * y + V[2, 2] * z + V[2, 3]
    idx = torch.nonzero(tz_all.detach() > c(NEAR_Z)).squeeze(1)
    x, y, z = x[idx], y[idx], z[idx]
    tx = V[0, 0] * x + V[0, 1] * y + V[0, 2] * z + V[0, 3]
    ty = V[1, 0] * x + V[1, 1] * y + V[1, 2] * z + V[1, 3]
    tz = V[2, 0] * x + V[2, 1] * y + V[2, 2] * z + V[2, 3]

    # -- 2. clip space: P . p_view (P = perspective only, renderer.py:57) -------
    hx = P[0, 0] * tx + P[0, 1] * ty + P[0, 2] * tz + P[0, 3]
    hy = P[1, 0] * tx + P[1, 1] * ty + P[1, 2] * tz + P[1, 3]
    hw = P[3, 0] * tx + P[3, 1] * ty + P[3, 2] * tz + P[3, 3]
    pw = c(1.0) / (hw + c(1e-7))
    ndcx = hx * pw
    ndcy = hy * pw
    if means2D is not None:
        # gradient sink: dL/dmeans2D is per NDC unit (App. A.5). +0 is exact.
        ndcx = ndcx + means2D[idx, 0]
        ndcy = ndcy + means2D[idx, 1]

    # -- 3. 3D covariance from scale and (un-normalised) quaternion ------------
    s = scales[idx] * c(st.scale_modifier)
    q = rotations[idx]
    qr, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    two = c(2.0)
    one = c(1.0)
    R00 = one - two * (qy * qy + qz * qz)
    R01 = two * (qx * qy - qr * qz)
    R02 = two * (qx * qz + qr * qy)
    R10 = two * (qx * qy + qr * qz)
    R11 = one - two * (qx * qx + qz * qz)
    R12 = two * (qy * qz - qr * qx)
    R20 = two * (qx * qz - qr * qy)
    R21 = two * (qy * qz + qr * qx)
    R22 = one - two * (qx * qx + qy * qy)
    s0, s1, s2 = s[:, 0], s[:, 1], s[:, 2]
    M00, M01, M02 = R00 * s0, R01 * s1, R02 * s2
    M10, M11, M12 = R10 * s0, R11 * s1, R12 * s2
    M20, M21, M22 = R20 * s0, R21 * s1, R22 * s2
    S00 = M00 * M00 + M01 * M01 + M02 * M02
    S01 = M00 * M10 + M01 * M11 + M02 * M12
    S02 = M00 * M20 + M01 * M21 + M02 * M22
    S11 = M10 * M10 + M11 * M11 + M12 * M12
    S12 = M10 * M20 + M11 * M21 + M12 * M22
    S22 = M20 * M20 + M21 * M21 + M22 * M22

    # -- 4. EWA 2D covariance ---------------------------------------------------
    fx = c(float(W)) / (two * tanx)
    fy = c(float(H)) / (two * tany)
    limx = c(1.3) * tanx
    limy = c(1.3) * tany
    # enable_cov_grad gates only the *pose* gradient of this path; the Gaussian's
    # own mean still receives it.  (Values are bit-identical to tx,ty,tz.)
    Vc = V if st.enable_cov_grad else V.detach()
    ctx_ = Vc[0, 0] * x + Vc[0, 1] * y + Vc[0, 2] * z + Vc[0, 3]
    cty_ = Vc[1, 0] * x + Vc[1, 1] * y + Vc[1, 2] * z + Vc[1, 3]
    ctz_ = Vc[2, 0] * x + Vc[2, 1] * y + Vc[2, 2] * z + Vc[2, 3]
    txtz = ctx_ / ctz_
    tytz = cty_ / ctz_
    in_x = (txtz.detach() >= -limx) & (txtz.detach() <= limx)
    in_y = (tytz.detach() >= -limy) & (tytz.detach() <= limy)
    cx = torch.minimum(limx, torch.maximum(-limx, txtz.detach())) * ctz_.detach()
    cy = torch.minimum(limy, torch.maximum(-limy, tytz.detach())) * ctz_.detach()
    # value == clamp(t.x/t.z)*t.z; gradient: identity when inside, frozen when clamped (A.6 iv)
    ctx = torch.where(in_x, cx + (ctx_ - ctx_.detach()), cx)
    cty = torch.where(in_y, cy + (cty_ - cty_.detach()), cy)
    tz2 = ctz_ * ctz_
    J00 = fx / ctz_
    J02 = -(fx * ctx) / tz2
    J11 = fy / ctz_
    J12 = -(fy * cty) / tz2
    T00 = J00 * Vc[0, 0] + J02 * Vc[2, 0]
    T01 = J00 * Vc[0, 1] + J02 * Vc[2, 1]
    T02 = J00 * Vc[0, 2] + J02 * Vc[2, 2]
    T10 = J11 * Vc[1, 0] + J12 * Vc[2, 0]
    T11 = J11 * Vc[1, 1] + J12 * Vc[2, 1]
    T12 = J11 * Vc[1, 2] + J12 * Vc[2, 2]
    U00 = T00 * S00 + T01 * S01 + T02 * S02
    U01 = T00 * S01 + T01 * S11 + T02 * S12
    U02 = T00 * S02 + T01 * S12 + T02 * S22
    U10 = T10 * S00 + T11 * S01 + T12 * S02
    U11 = T10 * S01 + T11 * S11 + T12 * S12
    U12 = T10 * S02 + T11 * S12 + T12 * S22
    ca = U00 * T00 + U01 * T01 + U02 * T02 + c(LOWPASS)
    cb = U00 * T10 + U01 * T11 + U02 * T12
    cc = U10 * T10 + U11 * T11 + U12 * T12 + c(LOWPASS)

    # -- 5. conic ---------------------------------------------------------------
    det = ca * cc - cb * cb
    det_ok = det.detach() != 0
    det_safe = torch.where(det_ok, det, torch.ones_like(det))
    det_inv = one / det_safe
    conA, conB, conC = _ConicFromCov.apply(ca, cb, cc, det_safe.detach(), det_inv.detach())

    # -- 6./7./8. radius, pixel centre, tile rectangle (integer outputs) --------
    with torch.no_grad():
        mid = c(0.5) * (ca + cc)
        disc = torch.sqrt(torch.maximum(c(0.1), mid * mid - det))
        lam1 = mid + disc
        lam2 = mid - disc
        rad_f = torch.ceil(c(3.0) * torch.sqrt(torch.maximum(lam1, lam2)))
    px = ((ndcx + one) * c(float(W)) - one) * c(0.5)
    py = ((ndcy + one) * c(float(H)) - one) * c(0.5)
    with torch.no_grad():
        tile = c(float(TILE))

        def tclamp(v, hi):
            # (int) cast truncates toward zero, then clamp to [0, hi]
            return torch.clamp(torch.trunc(v), 0.0, float(hi)).to(torch.int32)

        rmin_x = tclamp((px - rad_f) / tile, gx)
        rmin_y = tclamp((py - rad_f) / tile, gy)
        rmax_x = tclamp((px + rad_f + c(float(TILE - 1))) / tile, gx)
        rmax_y = tclamp((py + rad_f + c(float(TILE - 1))) / tile, gy)
        tiles = (rmax_x - rmin_x) * (rmax_y - rmin_y)
        visible = det_ok & (tiles > 0)
        radii = torch.zeros(N, dtype=torch.int32)
        tiles_touched = torch.zeros(N, dtype=torch.int32)
        radii[idx] = torch.where(visible, rad_f.to(torch.int32), torch.zeros_like(tiles))
        tiles_touched[idx] = torch.where(visible, tiles, torch.zeros_like(tiles))

    # -- 9. colour --------------------------------------------------------------
    if colors_precomp is not None:
        rgb = colors_precomp[idx]
    else:
        Vs = V if st.enable_sh_grad else V.detach()
        # campos = -R^T T
        cpx = -(Vs[0, 0] * Vs[0, 3] + Vs[1, 0] * Vs[1, 3] + Vs[2, 0] * Vs[2, 3])
        cpy = -(Vs[0, 1] * Vs[0, 3] + Vs[1, 1] * Vs[1, 3] + Vs[2, 1] * Vs[2, 3])
        cpz = -(Vs[0, 2] * Vs[0, 3] + Vs[1, 2] * Vs[1, 3] + Vs[2, 2] * Vs[2, 3])
        d = torch.stack([x - cpx, y - cpy, z - cpz], dim=1)
        d = d / torch.sqrt((d * d).sum(dim=1, keepdim=True))
        raw = sh_to_rgb(int(st.sh_degree), shs[idx], d) + c(0.5)
        rgb = torch.clamp_min(raw, 0.0)  # gradient 0 where clamped (A.6 iii)

    return Preprocessed(idx=idx, visible=visible, radii=radii, tiles_touched=tiles_touched,
                        rect_min=torch.stack([rmin_x, rmin_y], 1),
                        rect_max=torch.stack([rmax_x, rmax_y], 1),
                        xy=torch.stack([px, py], 1), depth=tz,
                        conic=torch.stack([conA, conB, conC], 1),
                        opacity=opacities[idx].reshape(-1), rgb=rgb,
                        cov2d=torch.stack([ca, cb, cc], 1).detach())


class Binned(NamedTuple):
    keys_unsorted: torch.Tensor   # [D] int64  (tile << 32) | depth bits, emission order
    vals_unsorted: torch.Tensor   # [D] int32  global Gaussian id
    keys: torch.Tensor            # [D] int64  sorted
    vals: torch.Tensor            # [D] int32  sorted (global ids)
    local: torch.Tensor           # [D] int64  sorted, index into the Preprocessed subset
    ranges: torch.Tensor          # [tiles,2] int32 [start,end)
    point_offsets: torch.Tensor   # [N] int64 inclusive scan of tiles_touched


@torch.no_grad()
def bin_tiles(pp: Preprocessed, H: int, W: int) -> Binned:
    """scan, duplicateWithKeys (y outer, x inner), stable
    sort by the 64-bit key, per-tile [start,end)."""
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    cnt_m = torch.where(pp.visible, (pp.rect_max[:, 0] - pp.rect_min[:, 0]) *
                        (pp.rect_max[:, 1] - pp.rect_min[:, 1]),
                        torch.zeros_like(pp.rect_min[:, 0])).to(torch.int64)
    point_offsets = torch.cumsum(pp.tiles_touched.to(torch.int64), 0)
    M = cnt_m.shape[0]
    owner = torch.repeat_interleave(torch.arange(M), cnt_m)           # subset index per entry
    start = torch.cumsum(cnt_m, 0) - cnt_m
    loc = torch.arange(owner.shape[0]) - start[owner]
    wdt = (pp.rect_max[:, 0] - pp.rect_min[:, 0]).to(torch.int64)[owner]
    ty = pp.rect_min[:, 1].to(torch.int64)[owner] + loc // torch.clamp_min(wdt, 1)
    tx = pp.rect_min[:, 0].to(torch.int64)[owner] + loc % torch.clamp_min(wdt, 1)
    depth_bits = pp.depth.detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    keys = ((ty * gx + tx) << 32) | (depth_bits[owner] & 0xFFFFFFFF)
    vals = pp.idx[owner].to(torch.int32)
    order = torch.sort(keys, stable=True).indices
    skeys = keys[order]
    tiles_of = (skeys >> 32)
    ranges = torch.zeros(gx * gy, 2, dtype=torch.int32)
    if skeys.numel() > 0:
        t_ids = torch.arange(gx * gy)
        lo = torch.searchsorted(tiles_of, t_ids, right=False)
        hi = torch.searchsorted(tiles_of, t_ids, right=True)
        nonempty = hi > lo
        ranges[nonempty, 0] = lo[nonempty].to(torch.int32)
        ranges[nonempty, 1] = hi[nonempty].to(torch.int32)
    return Binned(keys_unsorted=keys, vals_unsorted=vals, keys=skeys, vals=vals[order],
                  local=owner[order], ranges=ranges, point_offsets=point_offsets)


class Blended(NamedTuple):
    color: torch.Tensor     # [3,H,W]
    depth: torch.Tensor     # [1,H,W]
    alpha: torch.Tensor     # [1,H,W]
    final_T: torch.Tensor   # [H,W]
    n_contrib: torch.Tensor  # [H,W] int32
    pairs: int              # sum over tiles of n_tile * 256 (work measure)


def blend(pp: Preprocessed, bn: Binned, bg: torch.Tensor, H: int, W: int) -> Blended:
    """SURVEY.md App. A.4 (front-to-back alpha blend), one tile at a time as an
    [n_tile, 256] matrix; exclusive cumprod gives T_i."""
    dt = pp.xy.dtype
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    Hp, Wp = gy * TILE, gx * TILE
    color = torch.zeros(3, Hp, Wp, dtype=dt)
    depth = torch.zeros(1, Hp, Wp, dtype=dt)
    alpha = torch.zeros(1, Hp, Wp, dtype=dt)
    final_T = torch.ones(Hp, Wp, dtype=dt)
    n_contrib = torch.zeros(Hp, Wp, dtype=torch.int32)
    col_tiles, dep_tiles, alp_tiles = {}, {}, {}
    yy, xx = torch.meshgrid(torch.arange(TILE), torch.arange(TILE), indexing="ij")
    pairs = 0
    bgc = bg.to(dt).reshape(3, 1)
    for t in range(gx * gy):
        lo, hi = int(bn.ranges[t, 0]), int(bn.ranges[t, 1])
        if hi <= lo:
            continue
        pairs += (hi - lo) * TILE * TILE
        ty, tx = t // gx, t % gx
        pixx = (tx * TILE + xx).reshape(1, -1).to(dt)
        pixy = (ty * TILE + yy).reshape(1, -1).to(dt)
        g = bn.local[lo:hi]
        dx = pp.xy[g, 0:1] - pixx
        dy = pp.xy[g, 1:2] - pixy
        A, B, C = pp.conic[g, 0:1], pp.conic[g, 1:2], pp.conic[g, 2:3]
        power = -0.5 * (A * dx * dx + C * dy * dy) - B * dx * dy
        a_raw = pp.opacity[g].reshape(-1, 1) * torch.exp(power)
        a_cl = a_raw + (torch.clamp_max(a_raw, ALPHA_MAX) - a_raw).detach()  # straight-through
        valid = (power.detach() <= 0) & (a_cl.detach() >= ALPHA_MIN)
        a_eff = torch.where(valid, a_cl, torch.zeros_like(a_cl))
        one_m = 1.0 - a_eff
        T_incl = torch.cumprod(one_m, dim=0)                         # T after Gaussian i
        T_excl = torch.cat([torch.ones(1, T_incl.shape[1], dtype=dt), T_incl[:-1]], 0)
        stopped = T_incl.detach() < T_STOP                            # monotone in i
        contrib = valid & ~stopped
        w = torch.where(contrib, a_eff * T_excl, torch.zeros_like(a_eff))   # [n,256]
        # T_final = product over contributing Gaussians only
        T_fin = torch.prod(torch.where(contrib, one_m, torch.ones_like(one_m)), dim=0)
        ctile = (pp.rgb[g].unsqueeze(2) * w.unsqueeze(1)).sum(0)                  # [3,256]
        dtile = (pp.depth[g].reshape(-1, 1) * w).sum(0, keepdim=True)
        atile = w.sum(0, keepdim=True)
        ctile = ctile + T_fin.unsqueeze(0) * bgc
        col_tiles[t] = ctile
        dep_tiles[t] = dtile
        alp_tiles[t] = atile
        with torch.no_grad():
            ys, xs = ty * TILE, tx * TILE
            final_T[ys:ys + TILE, xs:xs + TILE] = T_fin.reshape(TILE, TILE)
            pos = torch.arange(1, hi - lo + 1).reshape(-1, 1)
            last = torch.where(contrib, pos, torch.zeros_like(pos)).max(dim=0).values
            n_contrib[ys:ys + TILE, xs:xs + TILE] = last.reshape(TILE, TILE).to(torch.int32)
    # assemble differentiably
    rows_c, rows_d, rows_a = [], [], []
    bg_tile = bgc.expand(3, TILE * TILE)
    z1 = torch.zeros(1, TILE * TILE, dtype=dt)
    for ty in range(gy):
        rc, rd, ra = [], [], []
        for tx in range(gx):
            t = ty * gx + tx
            rc.append(col_tiles.get(t, bg_tile).reshape(3, TILE, TILE))
            rd.append(dep_tiles.get(t, z1).reshape(1, TILE, TILE))
            ra.append(alp_tiles.get(t, z1).reshape(1, TILE, TILE))
        rows_c.append(torch.cat(rc, 2))
        rows_d.append(torch.cat(rd, 2))
        rows_a.append(torch.cat(ra, 2))
    color = torch.cat(rows_c, 1)[:, :H, :W]
    depth = torch.cat(rows_d, 1)[:, :H, :W]
    alpha = torch.cat(rows_a, 1)[:, :H, :W]
    return Blended(color=color, depth=depth, alpha=alpha, final_T=final_T[:H, :W],
                   n_contrib=n_contrib[:H, :W], pairs=pairs)


class RasterOut(NamedTuple):
    color: torch.Tensor
    depth: torch.Tensor
    normal: torch.Tensor
    alpha: torch.Tensor
    radii: torch.Tensor
    extra: Optional[torch.Tensor]
    pp: Preprocessed
    bn: Binned
    bl: Blended


def rasterize(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
              viewmatrix, st: Settings) -> RasterOut:
    """Same argument meaning as ``GaussianRasterizer.forward`` at
    renderer.py:88-100 (``cov3Ds_precomp`` unsupported - RoDyGS never passes it);
    returns the 6-tuple of renderer.py:87 plus the intermediates."""
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide exactly one of either SHs or precomputed colors!")
    H, W = int(st.image_height), int(st.image_width)
    pp = preprocess(means3D, scales, rotations, opacities, shs, colors_precomp, viewmatrix, st, means2D)
    bn = bin_tiles(pp, H, W)
    bl = blend(pp, bn, st.bg, H, W)
    normal = torch.zeros(3, H, W, dtype=bl.color.dtype)
    return RasterOut(bl.color, bl.depth, normal, bl.alpha, pp.radii, None, pp, bn, bl)
