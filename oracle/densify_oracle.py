"""PyTorch-CPU restatement of the reference's densification bookkeeping, stage by stage
(clone -> cat, split -> cat -> prune parents, prune by opacity / size), with the Adam-state
surgery of every stage.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the reference's own
Python (``DynTrainer.densify_and_prune`` executed on CPU) through ``tests/golden/densify.npz``.

Follows
* ``/root/reference/src/trainer/rodygs.py:316-341``            (per-iteration statistics)
* ``/root/reference/src/trainer/rodygs_static.py:150-159``     (reset_opacity)
* ``/root/reference/src/trainer/rodygs_static.py:170-319``     (postfix, clone, split, densify_and_prune, prune_points)
* ``/root/reference/src/trainer/rodygs_dynamic.py:150-197``    (motion_coeff rides along)
* ``/root/reference/src/trainer/utils.py:15-95``               (replace / cat / prune of the Adam moments)
* ``/root/reference/src/utils/general_utils.py:36-37,92-115``  (inverse_sigmoid, build_rotation)

The only liberty: ``torch.normal(mean=0, std=stds)`` (:184) is written as ``noise * stds`` with the
unit-normal ``noise [2 S, 3]`` passed in, so that the CUDA path can be fed the same samples.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

PARAMS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "motion_coeff")


def build_rotation(r: torch.Tensor) -> torch.Tensor:
    """general_utils.py:92-115 (normalises the quaternion, (r, x, y, z) order)."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), dtype=r.dtype)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r_ * z)
    R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y)
    R[:, 2, 1] = 2 * (y * z + r_ * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def get_scaling(scaling: torch.Tensor) -> torch.Tensor:
    s = torch.exp(scaling)
    return s.repeat(1, 3) if s.shape[1] == 1 else s          # rodygs_static.py:82-87


def add_densification_stats(radii, means2D_grad, max_radii2D, grad_accum, denom):
    """rodygs.py:319-341 for the rows of the model being trained (already sliced).
    Returns the updated (max_radii2D [n], grad_accum [n,1], denom [n,1])."""
    vis = radii > 0
    g = torch.norm(means2D_grad[:, :2], dim=-1, keepdim=True)
    max_radii2D = max_radii2D.clone()
    grad_accum = grad_accum.clone()
    denom = denom.clone()
    max_radii2D[vis] = torch.max(max_radii2D[vis], radii[vis].to(max_radii2D.dtype))
    grad_accum[vis] += g[vis]
    denom[vis] += 1
    return max_radii2D, grad_accum, denom


def reset_opacity(opacity: torch.Tensor, cap: float = 0.01):
    """rodygs_static.py:150-159: new raw opacity; the caller zeroes the group's moments (utils.py:15-32)."""
    o = torch.min(torch.sigmoid(opacity), torch.ones_like(opacity) * cap)
    return torch.log(o / (1 - o))


def _cat(state, ext):
    """cat_tensors_to_optimizer (utils.py:34-69): parameters grow by `ext`, both moments by zeros."""
    out = {}
    for k, (p, m, v) in state.items():
        e = ext[k]
        out[k] = (torch.cat((p, e), 0), torch.cat((m, torch.zeros_like(e)), 0), torch.cat((v, torch.zeros_like(e)), 0))
    return out


def _prune(state, valid):
    """prune_optimizer (utils.py:72-95)."""
    return {k: (p[valid], m[valid], v[valid]) for k, (p, m, v) in state.items()}


def densify_and_prune(state: Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], time: torch.Tensor,
                      time_ind: torch.Tensor, grad_accum: torch.Tensor, denom: torch.Tensor, max_grad: float,
                      min_opacity: float, extent: float, max_screen_size: Optional[float], percent_dense: float,
                      noise_fn, n_split: int = 2):
    """state: name -> (param, exp_avg, exp_avg_sq) for the names in PARAMS (motion_coeff optional).
    noise_fn(S) -> unit normal [n_split * S, 3].  Returns (state, time, time_ind, counts) where every
    statistic of the new model is zero (densification_postfix)."""
    grads = grad_accum / denom                                            # :286-287
    grads[grads.isnan()] = 0.0

    # ---- densify_and_clone (:246-283) ----
    sel = torch.norm(grads, dim=-1) >= max_grad
    sel = torch.logical_and(sel, torch.max(get_scaling(state["scaling"][0]), dim=1).values <= percent_dense * extent)
    ext = {k: state[k][0][sel] for k in state}
    time = torch.cat((time, time[sel]))
    time_ind = torch.cat((time_ind, time_ind[sel]))
    state = _cat(state, ext)
    n_clone = int(sel.sum())
    max_radii2D = torch.zeros(state["xyz"][0].shape[0])                   # densification_postfix (:165-170)

    # ---- densify_and_split (:172-244) ----
    n_init = state["xyz"][0].shape[0]
    padded = torch.zeros(n_init)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = padded >= max_grad
    scal = get_scaling(state["scaling"][0])
    sel = torch.logical_and(sel, torch.max(scal, dim=1).values > percent_dense * extent)
    S = int(sel.sum())
    stds = scal[sel].repeat(n_split, 1)
    samples = noise_fn(S) * stds                                          # torch.normal(mean=0, std=stds)
    rots = build_rotation(state["rotation"][0][sel]).repeat(n_split, 1, 1)
    new_scaling = torch.log(scal[sel].repeat(n_split, 1) / (0.8 * n_split))
    if state["scaling"][0].shape[1] == 1:
        new_scaling = new_scaling[:, [0]]
    ext = {k: state[k][0][sel].repeat(n_split, *([1] * (state[k][0].dim() - 1))) for k in state}
    ext["xyz"] = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + state["xyz"][0][sel].repeat(n_split, 1)
    ext["scaling"] = new_scaling
    state = _cat(state, ext)
    max_radii2D = torch.zeros(state["xyz"][0].shape[0])
    prune_filter = torch.cat((sel, torch.zeros(n_split * S, dtype=torch.bool)))
    time = torch.cat((time, *[time[sel] for _ in range(n_split)]))
    time_ind = torch.cat((time_ind, *[time_ind[sel] for _ in range(n_split)]))
    valid = ~prune_filter
    state = _prune(state, valid)
    max_radii2D, time, time_ind = max_radii2D[valid], time[valid], time_ind[valid]

    # ---- prune (:292-301) ----
    prune_mask = (torch.sigmoid(state["opacity"][0]) < min_opacity).squeeze(-1)
    if max_screen_size:
        big_vs = max_radii2D > max_screen_size
        big_ws = get_scaling(state["scaling"][0]).max(dim=1).values > 0.1 * extent
        prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_vs), big_ws)
    valid = ~prune_mask
    state = _prune(state, valid)
    time, time_ind = time[valid], time_ind[valid]
    return state, time, time_ind, {"clones": n_clone, "split": S, "rows": state["xyz"][0].shape[0]}
