"""Motion regularisers of the dynamic model, same class names, constructor arguments and
`forward(model, ...)` meaning as /root/reference/src/trainer/losses.py:

    RigidityLoss                 :185-361   modes "surface" / "distance_preserving" (every train config); "coeff" raises
    MotionL1Loss                 :364-367
    MotionSparsityLoss           :370-379
    MotionBasisRegularizaiton    :382-525   (the reference's spelling), degree 0 or negative

Each is ONE C-ABI call that returns the value together with its gradient (rodygs_b200/csrc/motion_reg.cu, knn.cu,
rigidity.cu); autograd only scales that gradient.  `model` is anything with the reference's attributes
(`_xyz`, `_motion_coeff [N,1,B]`, `unique_times`, `get_total_motion_table()`, `get_motion_for_times(...)`).
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr


def _cuda_f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("rodygs_b200 runs on CUDA tensors only (no CPU fallback)")
    return t.detach().float().contiguous()


def _ws(nbytes: int, dev) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


# ---- coefficients ------------------------------------------------------------------------------------------------

def motion_coeff_reg_(coeff: torch.Tensor, w_l1: float, w_sparsity: float, d_coeff: Optional[torch.Tensor],
                      accumulate: bool = True) -> torch.Tensor:
    """In-place form for the flat-buffer trainer: returns parts [2] = (mean|c|, sparsity) and adds
    w_l1 * dL1 + w_sparsity * dSparsity to d_coeff (same shape as coeff)."""
    lib = _lib.load()
    n, nb = coeff.numel() // coeff.shape[-1], coeff.shape[-1]
    parts = torch.empty(2, dtype=torch.float32, device=coeff.device)
    ws = _ws(16, coeff.device)
    check(lib.rdg_motion_coeff_reg(n, nb, ptr(coeff), float(w_l1), float(w_sparsity), ptr(parts), ptr(d_coeff),
                                   int(accumulate), ptr(ws), _lib.stream_ptr()))
    return parts


class _CoeffRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeff, which):
        c = _cuda_f32(coeff)
        grad = torch.empty_like(c) if coeff.requires_grad else None
        parts = motion_coeff_reg_(c, 1.0 if which == 0 else 0.0, 1.0 if which == 1 else 0.0, grad, accumulate=False)
        ctx.grad = grad
        ctx.shape = coeff.shape
        return parts[which].clone()

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g).view(ctx.shape), None


def motion_l1_loss(coeff: torch.Tensor) -> torch.Tensor:
    return _CoeffRegFn.apply(coeff, 0)


def motion_sparsity_loss(coeff: torch.Tensor) -> torch.Tensor:
    return _CoeffRegFn.apply(coeff, 1)


class MotionL1Loss(nn.Module):
    def forward(self, model, **kwargs):
        return motion_l1_loss(model._motion_coeff)


class MotionSparsityLoss(nn.Module):
    def forward(self, model, **kwargs):
        return motion_sparsity_loss(model._motion_coeff)


# ---- basis table ---------------------------------------------------------------------------------------------------

# losses.py:387-482 (16 bases); rebuilt from the closed forms the reference's literals come from would lose digits,
# so the published weights are kept to the digits the parity tests need (float32).
COEFF_BANK = {
    "gaussian": [2.368737348178644, 2.3218332060968687, 2.186620166400238, 1.9785357455909518, 1.7200563444604107,
                 1.4367118264767467, 1.1529882480025957, 0.8890134170352768, 0.6585973377702478, 0.4687700396753248,
                 0.3205737399288996, 0.2106319563365025, 0.13296850925636292, 0.08064947764026723,
                 0.04699834214974086, 0.026314295000921823],
    "sigmoid": [0.0, 0.006057306357564347, 0.019407599012746118, 0.04848852855754725, 0.11024831053568876,
                0.23462085565239668, 0.4602813915432914, 0.8016437593070956, 1.1983562406929047, 1.539718608456709,
                1.7653791443476032, 1.889751689464311, 1.9515114714424528, 1.9805924009872535, 1.9939426936424351, 2.0],
    "laplacian": [3.0235547043507864, 2.475477220065594, 2.0267493286116927, 1.6593620041145454, 1.3585707032576908,
                  1.112303614987853, 0.910677176350366, 0.7455994104042655, 0.6104451667747834, 0.49979023110633275,
                  0.40919363229470634, 0.3350194107233597, 0.274290694437278, 0.22457022681891523,
                  0.18386255092234366, 0.15053392477948924],
    "cum_exponential": [0.24858106424723717, 0.45210202617930384, 0.6187308966091, 0.7551550771806206,
                        0.8668497492779882, 0.9582976122790642, 1.0331687900213073, 1.0944681257580495,
                        1.1446557770689725, 1.1857459506219796, 1.219387739359138, 1.246931306386802,
                        1.2694820717618154, 1.2879450768797849, 1.3030613069641026, 1.3154374294047362],
    "vanilla": [1.0] * 16,
}


def basis_reg_coeff(freq_div_mode: str) -> torch.Tensor:
    """losses.py:483-489: weights / max * 1.3, except "vanilla"."""
    if freq_div_mode not in COEFF_BANK:
        raise AssertionError(f"Invalid freq_div_mode : {freq_div_mode}")
    w = torch.tensor(COEFF_BANK[freq_div_mode])
    return w / w.max() * 1.3 if freq_div_mode != "vanilla" else w


def motion_basis_reg_(table: torch.Tensor, reg_coeff: torch.Tensor, transl_degree: int, rot_degree: int,
                      d_table: Optional[torch.Tensor], grad_scale: float = 1.0) -> torch.Tensor:
    """In-place form: parts [2] = (translation term, rotation term); d_table += grad_scale * gradient."""
    lib = _lib.load()
    T, nb = table.shape[0], table.shape[1]
    parts = torch.empty(2, dtype=torch.float32, device=table.device)
    ws = _ws(16, table.device)
    check(lib.rdg_motion_basis_reg(T, nb, ptr(table), ptr(reg_coeff), int(transl_degree), int(rot_degree),
                                   float(grad_scale), ptr(parts), ptr(d_table), ptr(ws), _lib.stream_ptr()))
    return parts


class _BasisRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, reg_coeff, td, rd):
        t = _cuda_f32(table)
        grad = torch.zeros_like(t) if table.requires_grad else None
        parts = motion_basis_reg_(t, reg_coeff, td, rd, grad)
        ctx.grad = grad
        ctx.shape = table.shape
        loss = parts.new_zeros(())
        if td >= 0:
            loss = loss + parts[0]
        if rd >= 0:
            loss = loss + parts[1]
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g).view(ctx.shape), None, None, None


class MotionBasisRegularizaiton(nn.Module):
    def __init__(self, transl_degree=0, rot_degree=0, freq_div_mode="vanilla"):
        super().__init__()
        if transl_degree > 0 or rot_degree > 0:
            raise NotImplementedError("only degree 0 (velocity, every reference config) or negative (off) is built")
        self.degree = {"transl": transl_degree, "rot": rot_degree}
        self.reg_coeff = basis_reg_coeff(freq_div_mode).cuda()

    def forward(self, model, **kwargs):
        table = model.get_total_motion_table()
        return _BasisRegFn.apply(table, self.reg_coeff.to(table.device), self.degree["transl"], self.degree["rot"])


# ---- rigidity --------------------------------------------------------------------------------------------------------

def knn_points(points: torch.Tensor, K: int):
    """pytorch3d.ops.knn_points(p[None], p[None], K) for one cloud: (dist2 [n, K], idx [n, K] int32), exact."""
    lib = _lib.load()
    p = _cuda_f32(points)
    n = p.shape[0]
    idx = torch.empty(n, K, dtype=torch.int32, device=p.device)
    d2 = torch.empty(n, K, dtype=torch.float32, device=p.device)
    nbytes = int(lib.rdg_knn_workspace_bytes(n))
    ws = _ws(nbytes, p.device)
    check(lib.rdg_knn(n, ptr(p), int(K), ptr(idx), ptr(d2), ptr(ws), nbytes, _lib.stream_ptr()))
    return d2, idx


class _RigidityFn(torch.autograd.Function):
    """(points, canon, coeff, table) of the sampled rows -> loss; gradients w.r.t. all four."""

    @staticmethod
    def forward(ctx, points, canon, coeff, table, frame_indices, K, surface, distance):
        lib = _lib.load()
        p = _cuda_f32(points)
        n = p.shape[0]
        dev = p.device
        d2, idx = knn_points(p, K)
        a = _lib.RdgRigidity()
        a.n, a.K, a.eps = n, int(K), 1e-6
        a.mode_surface, a.mode_distance = int(surface), int(distance)
        parts = torch.zeros(2, dtype=torch.float32, device=dev)
        d_points = torch.empty_like(p)
        keep = [p, d2, idx, parts, d_points]
        a.points, a.nn_idx, a.nn_dist2 = ptr(p), ptr(idx), ptr(d2)
        a.loss_parts, a.d_points = ptr(parts), ptr(d_points)
        d_canon = d_coeff = d_table = None
        n_frames = 0
        if distance:
            cn, cf, tb = _cuda_f32(canon), _cuda_f32(coeff).view(n, -1), _cuda_f32(table)
            fi = frame_indices.to(dev, torch.int32).contiguous()
            n_frames = fi.numel()
            d_canon, d_coeff, d_table = torch.empty_like(cn), torch.empty_like(cf), torch.zeros_like(tb)
            a.num_basis, a.n_frames = cf.shape[1], n_frames
            a.canon, a.coeff, a.table, a.frame_indices = ptr(cn), ptr(cf), ptr(tb), ptr(fi)
            a.d_canon, a.d_coeff, a.d_table = ptr(d_canon), ptr(d_coeff), ptr(d_table)
            keep += [cn, cf, tb, fi]
        nbytes = int(lib.rdg_rigidity_workspace_bytes(n, int(K), n_frames))
        ws = _ws(nbytes, dev)
        check(lib.rdg_rigidity(C.byref(a), ptr(ws), nbytes, _lib.stream_ptr()))
        ctx.grads = (d_points, d_canon, d_coeff, d_table)
        ctx.shapes = (points.shape, canon.shape, coeff.shape, table.shape)
        ctx.mark_non_differentiable(parts)
        return parts.sum(), parts

    @staticmethod
    def backward(ctx, g, _gparts):
        out = []
        for gr, shp in zip(ctx.grads, ctx.shapes):
            out.append(None if gr is None else (gr * g).view(shp))
        return (*out, None, None, None, None)


def rigidity_loss(xyz: torch.Tensor, motion_coeff: torch.Tensor, pred_translation: torch.Tensor, table: Optional[torch.Tensor],
                  indice: torch.Tensor, frame_indices: Optional[torch.Tensor], K: int = 8,
                  mode: Sequence[str] = ("distance_preserving", "surface"), return_parts: bool = False):
    """Functional form with the reference's random draws passed in (indice: rows sampled by `random.sample`,
    frame_indices: the `torch.randint` draw)."""
    for m in mode:
        assert m in ["coeff", "surface", "distance_preserving"], f"Invalid mode: {m}"
    if "coeff" in mode:
        raise NotImplementedError('RigidityLoss mode "coeff" is not built (no reference config uses it)')
    indice = indice.to(xyz.device)
    canon = xyz[indice]                                            # plumbing: the gathers stay in PyTorch
    points = canon + pred_translation[indice]
    coeff = motion_coeff[indice]
    distance = "distance_preserving" in mode
    if distance and (table is None or frame_indices is None):
        raise Exception("distance_preserving needs the motion table and the sampled frame indices")
    if table is None:
        table = xyz.new_zeros(1, coeff.shape[-1], 7)
    loss, parts = _RigidityFn.apply(points, canon, coeff, table, frame_indices, K, "surface" in mode, distance)
    return (loss, parts) if return_parts else loss


class RigidityLoss(nn.Module):
    def __init__(self, scale: float = 2, K: int = 8, sim_metric: str = "l2", dist_weight_lambda: float = 0.1,
                 color_sim: bool = True, dist_preserving_ratio=4, mode=["coeff"]):
        super().__init__()
        self.scale, self.K = scale, K
        self.sim_metric, self.dist_weight_lambda, self.color_sim = sim_metric, dist_weight_lambda, color_sim
        self.mode = list(mode)
        self.dist_preserving_ratio = dist_preserving_ratio
        for m in self.mode:
            assert m in ["coeff", "surface", "distance_preserving"], f"Invalid mode: {m}"

    def forward(self, model, pred_translation, **kwargs):
        xyz, coeff = model._xyz, model._motion_coeff
        n = len(xyz)
        scale = 1 / self.scale if self.scale > 1 else self.scale
        indice = torch.tensor(random.sample(range(n), int(n * scale)))           # losses.py:228-232
        frame_indices = table = None
        if "distance_preserving" in self.mode:
            total = model.unique_times
            frame_indices = torch.randint(0, len(total) - 1, (len(total) // self.dist_preserving_ratio,))   # :297-301
            table = model.get_total_motion_table()
        return rigidity_loss(xyz, coeff, pred_translation, table, indice, frame_indices, self.K, self.mode)
