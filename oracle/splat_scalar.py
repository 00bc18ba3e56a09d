"""Second restatement of the tile rasterizer as plain scalar loops.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``oracle/splat_oracle.py`` is vectorised PyTorch (per-tile
[n, 256] matrices, cumulative products, autograd for the backward pass); this file is deliberately structured the
other way round - one Gaussian, one tile entry, one pixel at a time, NumPy float32 scalars, an explicit sort of
(key, emission index) pairs, and the blend backward written out by hand from App. A.5 - so that a slip in one of
the two restatements shows up as a disagreement (``tests/test_oracle_scalar.py``, <= 64 Gaussians).  Like the
vectorised oracle it follows the spec recalled in SURVEY.md App. A: the rasterizer source itself is not in
``/root/reference`` (un-vendored ``diff_gauss_pose``, ``.gitmodules:1-4``), so parity with upstream is unpinned;
the call-site conventions are those of ``/root/reference/src/trainer/renderer.py:50-101``.

Every float32 expression that feeds an integer output keeps the operation order of App. A / the vectorised oracle
(one rounding per operation), so radii, tile rectangles, keys and ranges must agree bit for bit.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

TILE = 16
f32 = np.float32

# /root/reference/src/utils/sh_utils.py:24-41
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def _sh_colour(deg: int, sh: np.ndarray, d: np.ndarray) -> np.ndarray:
    """SH sum (float64; only feeds the 1e-4 outputs).  sh [16,3], d unit vector.  sh_utils.py:72-101."""
    x, y, z = (float(v) for v in d)
    sh = sh.astype(np.float64)
    res = C0 * sh[0]
    if deg > 0:
        res = res - C1 * y * sh[1] + C1 * z * sh[2] - C1 * x * sh[3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + C2[0] * xy * sh[4] + C2[1] * yz * sh[5] + C2[2] * (2 * zz - xx - yy) * sh[6] + C2[3] * xz * sh[7]
               + C2[4] * (xx - yy) * sh[8])
    if deg > 2:
        res = (res + C3[0] * y * (3 * xx - yy) * sh[9] + C3[1] * xy * z * sh[10] + C3[2] * y * (4 * zz - xx - yy) * sh[11]
               + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[12] + C3[4] * x * (4 * zz - xx - yy) * sh[13]
               + C3[5] * z * (xx - yy) * sh[14] + C3[6] * x * (xx - 3 * yy) * sh[15])
    return res


def preprocess(means3D, scales, rotations, opacities, shs, viewmatrix_t, projmatrix_t, H, W, tanfovx, tanfovy,
               scale_modifier=1.0, sh_degree=3, colors_precomp=None) -> List[Dict]:
    """App. A.2 per Gaussian.  Matrices arrive in glm storage (transposed).  Returns one dict per Gaussian:
    radius, rect, tiles, and for the visible ones xy, depth, conic, opacity, rgb."""
    V = np.asarray(viewmatrix_t, dtype=f32).T
    P = np.asarray(projmatrix_t, dtype=f32).T
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    one, two = f32(1.0), f32(2.0)
    fx = f32(W) / (two * f32(tanfovx))
    fy = f32(H) / (two * f32(tanfovy))
    limx, limy = f32(1.3) * f32(tanfovx), f32(1.3) * f32(tanfovy)
    R, Tv = V[:3, :3].astype(np.float64), V[:3, 3].astype(np.float64)
    campos = -(R.T @ Tv)
    out = []
    for i in range(means3D.shape[0]):
        g = {"radius": 0, "tiles": 0, "visible": False}
        out.append(g)
        x, y, z = (f32(v) for v in means3D[i])
        tz = V[2, 0] * x + V[2, 1] * y + V[2, 2] * z + V[2, 3]
        if not (tz > f32(0.2)):
            continue                                                  # step 1: near-plane cull
        tx = V[0, 0] * x + V[0, 1] * y + V[0, 2] * z + V[0, 3]
        ty = V[1, 0] * x + V[1, 1] * y + V[1, 2] * z + V[1, 3]
        hx = P[0, 0] * tx + P[0, 1] * ty + P[0, 2] * tz + P[0, 3]     # step 2: P . p_view
        hy = P[1, 0] * tx + P[1, 1] * ty + P[1, 2] * tz + P[1, 3]
        hw = P[3, 0] * tx + P[3, 1] * ty + P[3, 2] * tz + P[3, 3]
        pw = one / (hw + f32(1e-7))
        ndcx, ndcy = hx * pw, hy * pw
        s = [f32(v) * f32(scale_modifier) for v in scales[i]]         # step 3: Sigma = M M^T, M = R(q) diag(s)
        qr, qx, qy, qz = (f32(v) for v in rotations[i])
        Rm = [[one - two * (qy * qy + qz * qz), two * (qx * qy - qr * qz), two * (qx * qz + qr * qy)],
              [two * (qx * qy + qr * qz), one - two * (qx * qx + qz * qz), two * (qy * qz - qr * qx)],
              [two * (qx * qz - qr * qy), two * (qy * qz + qr * qx), one - two * (qx * qx + qy * qy)]]
        M = [[Rm[r][k] * s[k] for k in range(3)] for r in range(3)]

        def dot3(a, b):
            return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]

        S = [[dot3(M[r], M[c]) for c in range(3)] for r in range(3)]
        txtz, tytz = tx / tz, ty / tz                                 # step 4: EWA
        cx = min(limx, max(-limx, txtz)) * tz
        cy = min(limy, max(-limy, tytz)) * tz
        tz2 = tz * tz
        J00, J02 = fx / tz, -(fx * cx) / tz2
        J11, J12 = fy / tz, -(fy * cy) / tz2
        T0 = [J00 * V[0, k] + J02 * V[2, k] for k in range(3)]
        T1 = [J11 * V[1, k] + J12 * V[2, k] for k in range(3)]
        U0 = [T0[0] * S[0][k] + T0[1] * S[1][k] + T0[2] * S[2][k] for k in range(3)]
        U1 = [T1[0] * S[0][k] + T1[1] * S[1][k] + T1[2] * S[2][k] for k in range(3)]
        ca = U0[0] * T0[0] + U0[1] * T0[1] + U0[2] * T0[2] + f32(0.3)
        cb = U0[0] * T1[0] + U0[1] * T1[1] + U0[2] * T1[2]
        cc = U1[0] * T1[0] + U1[1] * T1[1] + U1[2] * T1[2] + f32(0.3)
        det = ca * cc - cb * cb                                       # step 5
        if det == 0:
            continue
        det_inv = one / det
        conic = (cc * det_inv, -cb * det_inv, ca * det_inv)
        mid = f32(0.5) * (ca + cc)                                    # step 6
        disc = np.sqrt(max(f32(0.1), mid * mid - det))
        radius_f = np.ceil(f32(3.0) * np.sqrt(max(mid + disc, mid - disc)))
        px = ((ndcx + one) * f32(W) - one) * f32(0.5)                 # step 7
        py = ((ndcy + one) * f32(H) - one) * f32(0.5)

        def tclamp(v, hi):                                            # step 8: (int) truncates toward zero
            return int(min(max(float(np.trunc(v)), 0.0), float(hi)))

        rmin = (tclamp((px - radius_f) / f32(TILE), gx), tclamp((py - radius_f) / f32(TILE), gy))
        rmax = (tclamp((px + radius_f + f32(TILE - 1)) / f32(TILE), gx), tclamp((py + radius_f + f32(TILE - 1)) / f32(TILE), gy))
        tiles = (rmax[0] - rmin[0]) * (rmax[1] - rmin[1])
        if tiles == 0:
            continue
        if colors_precomp is not None:                                # step 9
            rgb, clamped = np.asarray(colors_precomp[i], dtype=np.float64), np.zeros(3, bool)
        else:
            d = means3D[i].astype(np.float64) - campos
            d = d / math.sqrt(float(d @ d))
            raw = _sh_colour(sh_degree, shs[i], d) + 0.5
            clamped = raw < 0
            rgb = np.maximum(raw, 0.0)
        g.update(visible=True, radius=int(radius_f), tiles=tiles, rmin=rmin, rmax=rmax, xy=(px, py), depth=tz, conic=conic,
                 opacity=f32(opacities[i]), rgb=rgb, clamped=clamped)
    return out


def bin_tiles(pp: List[Dict], H: int, W: int):
    """App. A.3: emission (Gaussian ascending, y outer, x inner), stable sort by the 64-bit key, per-tile ranges."""
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    pairs = []
    for i, g in enumerate(pp):
        if not g["visible"]:
            continue
        bits = int(np.asarray(g["depth"], dtype=f32).view(np.uint32))
        for ty in range(g["rmin"][1], g["rmax"][1]):
            for tx in range(g["rmin"][0], g["rmax"][0]):
                pairs.append((((ty * gx + tx) << 32) | bits, len(pairs), i))
    pairs.sort(key=lambda p: (p[0], p[1]))                            # stable: ties keep emission order
    keys = [p[0] for p in pairs]
    vals = [p[2] for p in pairs]
    ranges = [[0, 0] for _ in range(gx * gy)]
    for pos, k in enumerate(keys):
        t = k >> 32
        if pos == 0 or (keys[pos - 1] >> 32) != t:
            ranges[t][0] = pos
        if pos + 1 == len(keys) or (keys[pos + 1] >> 32) != t:
            ranges[t][1] = pos + 1
    return keys, vals, ranges


def blend(pp: List[Dict], vals, ranges, bg, H: int, W: int):
    """App. A.4 per pixel.  Returns color [3,H,W], depth, alpha, final_T [H,W], n_contrib [H,W] (all float32 maths)."""
    gx = (W + TILE - 1) // TILE
    color = np.zeros((3, H, W), f32)
    depth = np.zeros((H, W), f32)
    alpha = np.zeros((H, W), f32)
    final_T = np.ones((H, W), f32)
    n_contrib = np.zeros((H, W), np.int32)
    for py in range(H):
        for px in range(W):
            lo, hi = ranges[(py // TILE) * gx + px // TILE]
            T, C, D, A, last = f32(1.0), np.zeros(3, f32), f32(0), f32(0), 0
            for k in range(lo, hi):
                g = pp[vals[k]]
                dx, dy = g["xy"][0] - f32(px), g["xy"][1] - f32(py)
                cA, cB, cC = g["conic"]
                power = f32(-0.5) * (cA * dx * dx + cC * dy * dy) - cB * dx * dy
                if power > 0:
                    continue
                a = min(f32(0.99), g["opacity"] * np.exp(power, dtype=f32))
                if a < f32(1.0 / 255.0):
                    continue
                test_T = T * (f32(1.0) - a)
                if test_T < f32(1e-4):
                    break                                             # this Gaussian is not blended
                w = a * T
                C += g["rgb"].astype(f32) * w
                D += g["depth"] * w
                A += w
                T = test_T
                last = k - lo + 1
            color[:, py, px] = C + T * np.asarray(bg, f32)
            depth[py, px], alpha[py, px], final_T[py, px], n_contrib[py, px] = D, A, T, last
    return color, depth, alpha, final_T, n_contrib


def blend_backward(pp: List[Dict], vals, ranges, bg, H: int, W: int, final_T, n_contrib, dL_dcolor, dL_ddepth, dL_dalpha):
    """App. A.5 per pixel, back to front, written out by hand (float64 accumulation).  Returns per Gaussian
    dL/dxy (pixel units), dL/dmean2D (per NDC unit: x W/2, y H/2), dL/dconic (A, B, C), dL/dopacity, dL/drgb, dL/ddepth."""
    gx = (W + TILE - 1) // TILE
    n = len(pp)
    out = {"xy": np.zeros((n, 2)), "mean2D": np.zeros((n, 2)), "conic": np.zeros((n, 3)), "opacity": np.zeros(n),
           "rgb": np.zeros((n, 3)), "depth": np.zeros(n)}
    bgv = np.asarray(bg, np.float64)
    for py in range(H):
        for px in range(W):
            lo, _ = ranges[(py // TILE) * gx + px // TILE]
            gC = dL_dcolor[:, py, px].astype(np.float64)
            gD, gA = float(dL_ddepth[py, px]), float(dL_dalpha[py, px])
            T = float(final_T[py, px])
            T_final = T
            rec_c, rec_d, rec_a = np.zeros(3), 0.0, 0.0               # accumulated behind
            last_a, last_c, last_d = 0.0, np.zeros(3), 0.0
            bg_dot = float(bgv @ gC)
            for k in range(lo + int(n_contrib[py, px]) - 1, lo - 1, -1):
                i = vals[k]
                g = pp[i]
                dx, dy = float(g["xy"][0]) - px, float(g["xy"][1]) - py
                cA, cB, cC = (float(v) for v in g["conic"])
                # the skip rules replayed with the forward's float32 arithmetic
                dxf, dyf = g["xy"][0] - f32(px), g["xy"][1] - f32(py)
                power32 = f32(-0.5) * (g["conic"][0] * dxf * dxf + g["conic"][2] * dyf * dyf) - g["conic"][1] * dxf * dyf
                if power32 > 0:
                    continue
                G32 = np.exp(power32, dtype=f32)
                a32 = min(f32(0.99), g["opacity"] * G32)
                if a32 < f32(1.0 / 255.0):
                    continue
                G, a, o = float(G32), float(a32), float(g["opacity"])
                T = T / (1.0 - a)
                # "accumulated behind" recursions of App. A.5
                rec_c = last_a * last_c + (1.0 - last_a) * rec_c
                rec_d = last_a * last_d + (1.0 - last_a) * rec_d
                rec_a = last_a * 1.0 + (1.0 - last_a) * rec_a
                last_a, last_c, last_d = a, g["rgb"].astype(np.float64), float(g["depth"])
                w = a * T
                out["rgb"][i] += w * gC
                out["depth"][i] += w * gD
                dL_da = float((last_c - rec_c) @ gC) * T + (last_d - rec_d) * T * gD + (1.0 - rec_a) * T * gA
                dL_da += -(T_final / (1.0 - a)) * bg_dot              # colour only: background term
                out["opacity"][i] += G * dL_da                        # straight-through clamp (A.6 i)
                dL_dG = o * dL_da
                gdx = G * dL_dG * (-cA * dx - cB * dy)                # d power / d dx = -A dx - B dy
                gdy = G * dL_dG * (-cC * dy - cB * dx)
                out["xy"][i] += (gdx, gdy)                            # d dx / d xy = +1
                out["mean2D"][i] += (gdx * 0.5 * W, gdy * 0.5 * H)
                out["conic"][i] += (-0.5 * G * dx * dx * dL_dG, -G * dx * dy * dL_dG, -0.5 * G * dy * dy * dL_dG)
    return out
