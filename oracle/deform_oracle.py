"""PyTorch-CPU restatement of the time-conditioned deformation and the
static||dynamic Gaussian assembly.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the
reference's own Python through ``tests/golden/deform_*.npz``.

Follows
* ``/root/reference/src/model/rodygs_dynamic.py:202-220`` (TimestepEmbedder.forward)
* ``/root/reference/src/model/rodygs_dynamic.py:243-327`` (MLPBasisNetwork: timenet + 16 heads)
* ``/root/reference/src/model/rodygs_dynamic.py:122-147`` (get_gaussian_deformation, inverse_motion)
* ``/root/reference/src/model/rodygs_static.py:82-105``   (activation getters)
* ``/root/reference/src/trainer/rodygs.py:68-113``        (get_GS_properties: concat order static -> dynamic)
"""
from __future__ import annotations

import math
from typing import Dict, NamedTuple

import torch
import torch.nn.functional as F


def time_embedding(t: torch.Tensor, multires: int = 26, log_sampling: bool = False) -> torch.Tensor:
    """t: scalar or [T] -> [.., 2*multires+1] = [t, sin(f0 pi t), cos(f0 pi t), sin(f1 pi t), ...]."""
    if log_sampling:
        freqs = 2.0 ** torch.linspace(0.0, multires - 1, multires)
    else:
        freqs = torch.linspace(1.0, 2.0 ** (multires - 1), multires)
    freqs = freqs * math.pi
    t = torch.as_tensor(t, dtype=torch.float32)
    parts = [t]
    for f in freqs:
        parts.append(torch.sin(t * f))
        parts.append(torch.cos(t * f))
    return torch.stack(parts, dim=-1)


def motion_basis(state: Dict[str, torch.Tensor], emb: torch.Tensor, num_basis: int = 16,
                 gelu: bool = True) -> torch.Tensor:
    """emb [..,53] -> B [..,num_basis,7] using the reference module's state_dict
    naming (timenet.{0,2,4}, basis_xyz.{k}.basis.{0,2})."""
    act = F.gelu if gelu else F.relu
    h = emb
    for li in (0, 2, 4):
        h = act(F.linear(h, state[f"timenet.{li}.weight"], state[f"timenet.{li}.bias"]))
    outs = []
    for k in range(num_basis):
        u = act(F.linear(h, state[f"basis_xyz.{k}.basis.0.weight"], state[f"basis_xyz.{k}.basis.0.bias"]))
        outs.append(F.linear(u, state[f"basis_xyz.{k}.basis.2.weight"], state[f"basis_xyz.{k}.basis.2.bias"]))
    return torch.stack(outs, dim=-2)


def gaussian_deformation(coeff: torch.Tensor, basis_t: torch.Tensor, table: torch.Tensor,
                         time_ind: torch.Tensor, spatial_lr_scale: float):
    """coeff [Nd,16] (the reference stores [Nd,1,16]); basis_t [16,7] = B(t);
    table [T,16,7] = B at every training time; time_ind [Nd] birth-frame index.
    Returns (scaled_translation [Nd,3], delta_rotation [Nd,4]) with the
    inverse-motion rule  c.B(t) - c.B(t_i)  (rodygs_dynamic.py:122-138)."""
    tot = coeff @ basis_t                                             # [Nd,7]
    birth = torch.bmm(coeff.unsqueeze(1), table[time_ind]).squeeze(1)  # [Nd,7]
    trans = tot[:, :3] - birth[:, :3]
    rot = tot[:, 3:] - birth[:, 3:]
    return trans * spatial_lr_scale, rot


class RawGaussians(NamedTuple):
    """Un-activated parameters of one model (rodygs_static.py:35-47)."""
    xyz: torch.Tensor            # [n,3]
    features_dc: torch.Tensor    # [n,1,3]
    features_rest: torch.Tensor  # [n,15,3]
    scaling: torch.Tensor        # [n,3] log-scale
    rotation: torch.Tensor       # [n,4] raw quaternion
    opacity: torch.Tensor        # [n,1] logit


def activate(g: RawGaussians):
    """rodygs_static.py:82-105."""
    return (g.xyz, torch.sigmoid(g.opacity), torch.exp(g.scaling), F.normalize(g.rotation),
            torch.cat((g.features_dc, g.features_rest), dim=1))


def assemble(static: RawGaussians, dynamic: RawGaussians, coeff, basis_t, table, time_ind,
             spatial_lr_scale: float, use_deform: bool = True):
    """rodygs.py:68-113: static first, dynamic second; the rotation delta is added
    to the *normalised* quaternion and the sum is not re-normalised."""
    sx, so, ss, sr, sf = activate(static)
    dx, do, ds, dr, df = activate(dynamic)
    if use_deform:
        dtrans, drot = gaussian_deformation(coeff, basis_t, table, time_ind, spatial_lr_scale)
        dx = dx + dtrans
        dr = dr + drot
    return (torch.cat((sx, dx), 0), torch.cat((so, do), 0), torch.cat((ss, ds), 0),
            torch.cat((sr, dr), 0), torch.cat((sf, df), 0))
