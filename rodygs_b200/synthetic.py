"""Seeded synthetic scenes and cameras of the shapes BASELINE.json names
(SURVEY.md §8d).  Everything is generated on the CPU from
`torch.Generator().manual_seed(...)` so that the oracle (CPU) and the CUDA path
(after `.to('cuda')`) see bit-identical inputs.  Camera conventions follow
/root/reference/src/utils/graphic_utils.py:43-63 (projection) and
/root/reference/src/data/utils.py:161-170 (world->view).
"""
from __future__ import annotations

import math
from typing import Dict, NamedTuple

import torch

CONFIGS = {
    # name: (N, H, W, T, views_per_step)   - 50 % of the Gaussians are dynamic
    "c1_cpu": (10_000, 256, 256, 8, 1),
    "c2_kubric": (300_000, 512, 512, 100, 1),
    "c3_nvidia": (1_000_000, 540, 960, 24, 8),
    "c4_iphone": (2_000_000, 1080, 1920, 100, 8),
    "c5_infer": (6_000_000, 1080, 1920, 100, 1),
}
FOCAL_OVER_W = 1.2
ZNEAR, ZFAR = 0.01, 100.0


class Camera(NamedTuple):
    height: int
    width: int
    FoVx: float
    FoVy: float
    tanfovx: float
    tanfovy: float
    world_view_transform: torch.Tensor   # V  [4,4] (mathematical, row-major)
    projection_matrix: torch.Tensor      # P  [4,4]
    time_index: int
    time: float


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """graphic_utils.py:43-63 (z_sign = +1)."""
    t, r = math.tan(fovy / 2) * znear, math.tan(fovx / 2) * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * r)
    P[1, 1] = 2.0 * znear / (2 * t)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(view: int, n_views: int, height: int, width: int, n_times: int) -> Camera:
    focal = FOCAL_OVER_W * width
    fovx = 2 * math.atan(width / (2 * focal))
    fovy = 2 * math.atan(height / (2 * focal))
    th = 2 * math.pi * view / max(n_views, 1)
    c = torch.tensor([0.5 * math.cos(th), 0.5 * math.sin(th), 0.0], dtype=torch.float64)
    target = torch.tensor([0.0, 0.0, 4.0], dtype=torch.float64)
    z = target - c
    z = z / z.norm()
    x = torch.linalg.cross(torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64), z)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    R = torch.stack([x, y, z])              # world->view rotation
    V = torch.eye(4, dtype=torch.float64)
    V[:3, :3] = R
    V[:3, 3] = -R @ c
    t_idx = int(round(view / max(n_views, 1) * n_times)) % n_times
    return Camera(height, width, fovx, fovy, math.tan(fovx * 0.5), math.tan(fovy * 0.5), V.float(),
                  projection_matrix(ZNEAR, ZFAR, fovx, fovy), t_idx, t_idx / n_times)


def make_scene(n: int, height: int, width: int, n_times: int, seed: int = 0, dynamic_fraction: float = 0.5,
               radius_px: float = 4.0, num_basis: int = 16) -> Dict[str, torch.Tensor]:
    """Raw (un-activated) parameters of a static and a dynamic model + the motion inputs."""
    g = torch.Generator().manual_seed(seed)
    n_dyn = int(n * dynamic_fraction)
    n_sta = n - n_dyn
    focal = FOCAL_OVER_W * width
    tanx, tany = width / (2 * focal), height / (2 * focal)

    def rnd(*shape):
        return torch.randn(*shape, generator=g)

    def uni(*shape):
        return torch.rand(*shape, generator=g)

    z = 2.0 + 6.0 * uni(n)
    x = (2 * uni(n) - 1) * 1.05 * tanx * z
    y = (2 * uni(n) - 1) * 1.05 * tany * z
    behind = uni(n) < 0.01                      # 1 % behind the near plane to exercise culling
    z = torch.where(behind, 0.19 * uni(n), z)
    xyz = torch.stack([x, y, z], 1)
    s0 = radius_px * 5.0 / (3.0 * focal)        # 3 sigma ~ radius_px at the median depth 5
    scaling = math.log(s0) + 0.4 * rnd(n, 3)
    rotation = rnd(n, 4)
    opacity = -1.0 + 1.5 * rnd(n, 1)
    f_dc = rnd(n, 1, 3)
    f_rest = 0.1 * rnd(n, 15, 3)
    perm = torch.randperm(n, generator=g)
    si, di = perm[:n_sta], perm[n_sta:]
    out = {}
    for name, idx in (("static", si), ("dynamic", di)):
        out[name] = {"xyz": xyz[idx].contiguous(), "features_dc": f_dc[idx].contiguous(),
                     "features_rest": f_rest[idx].contiguous(), "scaling": scaling[idx].contiguous(),
                     "rotation": rotation[idx].contiguous(), "opacity": opacity[idx].contiguous()}
    out["motion_coeff"] = 0.1 * rnd(n_dyn, 1, num_basis)
    out["time_ind"] = torch.randint(0, n_times, (n_dyn,), generator=g, dtype=torch.int32)
    out["table"] = 0.05 * rnd(n_times, num_basis, 7)
    out["spatial_lr_scale"] = 1.0
    return out


def to_device(tree, device):
    if isinstance(tree, torch.Tensor):
        return tree.to(device)
    if isinstance(tree, dict):
        return {k: to_device(v, device) for k, v in tree.items()}
    return tree


def algorithmic_bytes(n: int, visible: int, n_dyn: int, dups: int, pixels: int, sh_coeffs: int = 16,
                      forward_only: bool = False, with_loss: bool = True) -> int:
    """SURVEY.md §8d / BASELINE.md §3.3."""
    if forward_only:
        return 60 * n + (60 + 12 * sh_coeffs) * visible + 68 * n_dyn + 88 * dups + 28 * pixels
    per_px = 120 if with_loss else 120 - 48
    return 160 * n + (180 + 36 * sh_coeffs) * visible + 200 * n_dyn + 132 * dups + per_px * pixels
