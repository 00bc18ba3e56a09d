#!/usr/bin/env python
"""Freeze the rasterizer oracle (oracle/splat_oracle.py) on one fixed scene -> tests/golden/raster_small.npz.

The rasterizer is the one part of the hot path whose reference source is absent (un-vendored ``diff_gauss_pose``,
/root/reference/.gitmodules:1-4), so there is no reference output to generate; what is frozen here are the
ORACLE's outputs, so that an edit to the oracle cannot silently move the target of rows a7-a10:
tests/test_oracle_golden.py::test_rasterizer_oracle_is_frozen re-runs the oracle against this file on the CPU,
tests/test_gpu_golden.py compares the CUDA path with it directly.

    python tests/golden/make_golden_raster.py        (from the repo root; CPU only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle import splat_oracle as so  # noqa: E402

N, H, W, DEG, SEED = 600, 72, 104, 3, 20261017


def build():
    torch.set_num_threads(1)
    sc, cam = helpers.small_scene(N, H, W, 5, seed=SEED, radius_px=5.0)
    acts = helpers.activated_concat(sc, cam)
    xyz, op, scl, rot, feat = [t.detach().clone().requires_grad_(True) for t in acts]
    bg = torch.tensor([0.25, 0.05, 0.4])
    vm = cam.world_view_transform.t().contiguous().requires_grad_(True)
    m2 = torch.zeros(N, 3, requires_grad=True)
    st = helpers.oracle_settings(cam, bg, DEG)
    out = so.rasterize(xyz, m2, feat, None, op, scl, rot, vm, st)
    g = torch.Generator().manual_seed(SEED + 1)
    gc, gd, ga = torch.randn(3, H, W, generator=g), 0.3 * torch.randn(1, H, W, generator=g), 0.2 * torch.randn(1, H, W, generator=g)
    ((out.color * gc).sum() + (out.depth * gd).sum() + (out.alpha * ga).sum()).backward()
    d = dict(
        means3D=xyz, opacities=op, scales=scl, rotations=rot, shs=feat, viewmatrix=vm, bg=bg,
        projmatrix=cam.projection_matrix.t().contiguous(), tanfov=torch.tensor([cam.tanfovx, cam.tanfovy], dtype=torch.float64),
        dL_dcolor=gc, dL_ddepth=gd, dL_dalpha=ga,
        radii=out.radii, tiles_touched=out.pp.tiles_touched, keys=out.bn.keys, vals=out.bn.vals, ranges=out.bn.ranges,
        color=out.color, depth=out.depth, alpha=out.alpha, final_T=out.bl.final_T,
        g_means3D=xyz.grad, g_means2D=m2.grad, g_shs=feat.grad, g_opacities=op.grad, g_scales=scl.grad,
        g_rotations=rot.grad, g_viewmatrix=vm.grad)
    return {k: v.detach().numpy() for k, v in d.items()}, dict(n=N, H=H, W=W, deg=DEG)


if __name__ == "__main__":
    data, meta = build()
    path = os.path.join(ROOT, "tests", "golden", "raster_small.npz")
    np.savez_compressed(path, meta=np.array([meta["n"], meta["H"], meta["W"], meta["deg"]]), **data)
    print(path, os.path.getsize(path), "bytes; duplicates:", data["keys"].shape[0], "visible:", int((data["radii"] > 0).sum()))
