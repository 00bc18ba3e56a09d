"""PyTorch-CPU restatement of the reference's motion regularisers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows
* ``/root/reference/src/trainer/losses.py:185-361``  RigidityLoss (modes surface / coeff / distance_preserving)
* ``/root/reference/src/trainer/losses.py:364-379``  MotionL1Loss, MotionSparsityLoss
* ``/root/reference/src/trainer/losses.py:382-525``  MotionBasisRegularizaiton
* ``/root/reference/src/utils/loss_utils.py:206-250`` get_outnorm, CharbonnierLoss (eps 1e-6, out_norm "bc")
* ``/root/reference/src/utils/graphic_utils.py:76-102`` quaternion_to_matrix

Parity status: **pinned** for everything after the K-nearest-neighbour search, against the reference's own
classes executed on CPU (``tests/golden/motion.npz``, generator ``tests/golden/make_golden_motion.py``).
The search itself is ``pytorch3d.ops.knn_points`` - a third-party dependency installed from the un-vendored
``thirdparty/pytorch3d`` (README.md:35, no pinned version) that is absent here; ``knn_points`` below restates its
documented contract (squared Euclidean distances, the K smallest in ascending order, the query itself included
when the two clouds are the same) by exhaustive search, and the generator binds the reference to this
restatement.  Ties are broken towards the lower index (pytorch3d leaves the order of exact ties unspecified).

Every function takes the random draws of the reference (``random.sample`` of the Gaussians, ``torch.randint`` of
the time indices) as arguments so that the CUDA path can be fed the same ones.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn.functional as F

# MotionBasisRegularizaiton.coeff_bank (losses.py:387-482) is reproduced by the product in
# rodygs_b200/motion_reg.py; the oracle receives the already normalised 16 weights as an argument.


def knn_points(points: torch.Tensor, K: int, chunk: int = 2048) -> Tuple[torch.Tensor, torch.Tensor]:
    """pytorch3d.ops.knn_points(p[None], p[None], K) -> (dists [n, K] squared, idx [n, K] int64).
    The distance is accumulated as (dx*dx + dy*dy) + dz*dz in the input precision, one rounding per operation."""
    n = points.shape[0]
    idx = torch.empty(n, K, dtype=torch.int64)
    dist = torch.empty(n, K, dtype=points.dtype)
    ar = torch.arange(n)
    with torch.no_grad():
        for s in range(0, n, chunk):
            q = points[s:s + chunk]
            dx = q[:, None, 0] - points[None, :, 0]
            dy = q[:, None, 1] - points[None, :, 1]
            dz = q[:, None, 2] - points[None, :, 2]
            d = (dx * dx + dy * dy) + dz * dz
            # stable sort on the distance => equal distances keep ascending index order
            order = torch.sort(d, dim=1, stable=True).indices[:, :K]
            idx[s:s + chunk] = order
    # differentiable distances, the way pytorch3d's autograd returns them
    nb = points[idx]                                    # [n, K, 3]
    dx = points[:, None, 0] - nb[..., 0]
    dy = points[:, None, 1] - nb[..., 1]
    dz = points[:, None, 2] - nb[..., 2]
    dist = (dx * dx + dy * dy) + dz * dz
    del ar
    return dist, idx


def motion_l1(coeff: torch.Tensor) -> torch.Tensor:
    """MotionL1Loss (losses.py:364-367)."""
    return coeff.abs().mean()


def motion_sparsity(coeff: torch.Tensor) -> torch.Tensor:
    """MotionSparsityLoss (losses.py:370-379); coeff [N, 1, B]."""
    a = torch.abs(coeff)
    mx, _ = torch.max(a, dim=2)
    return (a / (mx[..., None] + 1e-7)).mean()


def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """graphic_utils.py:76-102 (real part first; scaled by 2 / |q|^2, no normalisation beforehand)."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def motion_basis_reg(table: torch.Tensor, reg_coeff: torch.Tensor, transl_degree: int = 0, rot_degree: int = 0):
    """MotionBasisRegularizaiton.forward (losses.py:499-525); table [T, B, 7].
    `derivate_motion` (:493-497) never passes is_rot, so the rotation branch is a plain finite difference of
    the rotation MATRICES, and the penalty is || I - (R[t+1] - R[t]) ||_F - kept as the reference computes it."""
    transl, rot = table[..., :3], table[..., 3:]
    R = quaternion_to_matrix(rot.reshape(-1, 4)).reshape(*table.shape[:-1], 3, 3)
    for _ in range(transl_degree + 1):
        transl = transl[1:] - transl[:-1]
    for _ in range(rot_degree + 1):
        R = R[1:] - R[:-1]
    t_norm = (torch.norm(transl, dim=-1) * reg_coeff[None]).mean()
    r_norm = (torch.norm(torch.eye(3, dtype=table.dtype)[None, None] - R, dim=(-1, -2)) * reg_coeff[None]).mean()
    if transl_degree < 0:
        t_norm = 0
    if rot_degree < 0:
        r_norm = 0
    return t_norm + r_norm


def rigidity(xyz: torch.Tensor, coeff: torch.Tensor, features_dc: torch.Tensor, pred_translation: torch.Tensor,
             table: torch.Tensor, indice: torch.Tensor, time_indices: torch.Tensor, K: int = 8,
             mode: Sequence[str] = ("distance_preserving", "surface"), sim_metric: str = "l2",
             dist_weight_lambda: float = 0.1, color_sim: bool = True, eps: float = 1e-6):
    """RigidityLoss.forward (losses.py:216-361).  xyz [N,3], coeff [N,1,B], features_dc [N,1,3],
    pred_translation [N,3], table [T,B,7]; indice = the `random.sample` draw (:228-232),
    time_indices = the `torch.randint` draw (:297-301).  Returns (loss, parts dict)."""
    gp = xyz + pred_translation
    tp, tc, tcol = gp[indice], coeff[indice], features_dc[indice]
    n = tp.shape[0]
    dd, nn = knn_points(tp, K)                                       # :238-239
    parts = {}
    loss = torch.zeros((), dtype=xyz.dtype)
    if "surface" in mode:                                            # :243-254
        nnp = tp[nn]                                                 # [n, K, 3]
        s = F.pairwise_distance(tp, nnp.mean(dim=1), p=2).mean()
        parts["surface"] = s
        loss = loss + s
    if "coeff" in mode:                                              # :256-297
        cn = tc[nn]                                                  # [n, K, 1, B]
        coln = tcol.view(-1, 3)[nn]                                  # [n, K, 3]
        # target_colors is [n,1,3]: F.pairwise_distance(target_colors[None], color_nn[None]) broadcasts to [1,n,K]
        cd = F.pairwise_distance(tcol[None], coln[None], p=2)[0]
        dw = torch.exp(-dist_weight_lambda * dd ** 2)
        cw = torch.exp(-dist_weight_lambda * cd ** 2)
        sc = tc[:, None]                                             # [n, 1, 1, B]
        if sim_metric == "cosine":
            sim = F.cosine_similarity(sc, cn, dim=2)
        elif sim_metric == "l2":
            sim = F.pairwise_distance(sc, cn, p=2)
        elif sim_metric == "l1":
            sim = F.pairwise_distance(sc, cn, p=1)
        else:
            raise ValueError("Invalid similarity metric")
        sim = (cw * dw if color_sim else dw) * sim.squeeze()
        parts["coeff"] = sim.mean()
        loss = loss + sim.mean()
    if "distance_preserving" in mode:                                # :299-358
        Ts = time_indices.shape[0]
        B3 = table[time_indices][..., :3]                            # [Ts, B, 3]
        tr = (tc[:, None] @ B3).squeeze(2)                           # [n, Ts, 3]
        canon = xyz[indice]
        loc = canon[:, None] + tr                                    # [n, Ts, 3]
        loc_nn = loc[nn]                                             # [n, K, Ts, 3]
        diff = loc_nn.permute(2, 0, 1, 3) - loc.transpose(0, 1)[:, :, None]     # [Ts, n, K, 3]
        x = torch.norm(diff, dim=-1)                                 # [Ts, n, K]
        # :353-356 reinterprets the contiguous [1, Ts, n, K] block as [n*K, Ts, 1] (a memory view, not a
        # transpose) and pairs row r with the squared neighbour distance number r
        xv = x.contiguous().view(-1, Ts, 1)
        yv = dd[None].reshape(-1, 1, 1)
        norm = 1.0 / (xv.shape[0] * xv.shape[1])                     # get_outnorm "bc"
        dp = torch.sum(torch.sqrt((xv - yv).pow(2) + eps ** 2)) * norm
        parts["distance_preserving"] = dp
        loss = loss + dp
    return loss, parts
