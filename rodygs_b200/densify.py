"""Densification on the GPU-resident SoA model: per-iteration statistics,
densify_and_prune as one stream compaction, opacity reset.  Host mirror of

* /root/reference/src/trainer/rodygs.py:316-362            (statistics, scheduling)
* /root/reference/src/trainer/rodygs_static.py:150-319     (reset_opacity, clone / split / prune)
* /root/reference/src/trainer/rodygs_dynamic.py:150-197    (motion_coeff, gaussian_to_time(_ind) ride along)
* /root/reference/src/trainer/utils.py:15-95               (Adam moments follow their rows)

over the C ABI (rdg_densify_stats / _plan / _apply, rdg_reset_opacity).  No CPU fallback.

Data-parallel runs (one camera-time view per rank) must densify identically on every rank:
`DensifyStats.all_reduce()` sums the gradient statistics and the visibility counts and takes the
maximum of the radii over the ranks - the same numbers a single process would have gathered by
visiting the views one after the other - and the split noise comes from a generator seeded alike.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr

# name -> gather mode of the parameter itself (its Adam moments always use MOMENT)
MODE_COPY, MODE_MOMENT, MODE_XYZ, MODE_SCALING = 0, 1, 2, 3
PARAM_MODES = {"xyz": MODE_XYZ, "scaling": MODE_SCALING}


class DensifyStats:
    """max_radii2D / xyz_gradient_accum / denom of one model (rodygs_static.py:165-170)."""

    def __init__(self, n: int, device="cuda"):
        f32 = dict(dtype=torch.float32, device=device)
        self.max_radii2D = torch.zeros(n, **f32)
        self.grad_accum = torch.zeros(n, **f32)
        self.denom = torch.zeros(n, **f32)

    @property
    def n(self) -> int:
        return self.max_radii2D.shape[0]

    def add(self, radii: torch.Tensor, means2D_grad: torch.Tensor, offset: int = 0):
        """rodygs.py:319-341: `radii` [N] int32 and `means2D_grad` [N,3] of the WHOLE concatenated scene;
        this model's rows are offset .. offset + n (static first, dynamic second, rodygs.py:87-100)."""
        _lib.require_cuda(radii, means2D_grad)
        n = self.n
        if radii.dtype != torch.int32 or means2D_grad.dtype != torch.float32:
            raise RuntimeError("radii must be int32 and means2D_grad float32")
        if offset < 0 or offset + n > radii.shape[0] or means2D_grad.shape[0] != radii.shape[0]:
            raise RuntimeError("statistics slice out of range")
        r = radii[offset:offset + n]
        g = means2D_grad[offset:offset + n]
        if not (r.is_contiguous() and g.is_contiguous() and g.shape[1] == 3):
            raise RuntimeError("radii / means2D_grad must be contiguous [N] / [N,3]")
        check(_lib.load().rdg_densify_stats(n, ptr(r), ptr(g), ptr(self.max_radii2D), ptr(self.grad_accum), ptr(self.denom),
                                            _lib.stream_ptr()))

    def all_reduce(self, group=None):
        """Make the statistics rank-consistent (sum / sum / max); a no-op for one process."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad_accum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.denom, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, group=group)


def _rows(t: torch.Tensor) -> Tuple[int, int]:
    n = t.shape[0]
    return n, (t.numel() // n if n else 0)


def densify_and_prune(params: Dict[str, torch.Tensor], moments: Optional[Dict[str, Tuple[torch.Tensor, torch.Tensor]]],
                      extras: Optional[Dict[str, torch.Tensor]], stats: DensifyStats, grad_threshold: float,
                      min_opacity: float, extent: float, max_screen_size: Optional[float], percent_dense: float = 0.01,
                      noise: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None):
    """ThreeDGSTrainer / DynTrainer.densify_and_prune(max_grad, min_opacity, extent, max_screen_size).

    params   name -> [n, ...] fp32 (must hold xyz, scaling, rotation, opacity; f_dc, f_rest, motion_coeff ... ride along)
    moments  name -> (exp_avg, exp_avg_sq) of the same shapes (None: no optimiser state)
    extras   name -> [n] 32-bit tensors copied with their rows (gaussian_to_time, gaussian_to_time_ind as int32)
    noise    [2 S, 3] unit normal samples for the split (S = number of split-selected rows; row c * S + k for copy c of
             the k-th selected row, the layout of the reference's torch.normal call).  Drawn from `generator` when None.
    Returns (new_params, new_moments, new_extras, new_stats, info); one host read of four counters
    (the reference reads the masks several times per stage)."""
    lib = _lib.load()
    stream = _lib.stream_ptr()
    xyz, scaling, rotation, opacity = params["xyz"], params["scaling"], params["rotation"], params["opacity"]
    _lib.require_cuda(*params.values())
    n = xyz.shape[0]
    dev = xyz.device
    if stats.n != n:
        raise RuntimeError("statistics do not match the model")
    sw = scaling.shape[1]
    for t in list(params.values()) + ([x for mv in moments.values() for x in mv] if moments else []) + list((extras or {}).values()):
        if t.shape[0] != n or not t.is_contiguous() or t.element_size() != 4:
            raise RuntimeError("every field must be a contiguous 32-bit tensor with n rows")
    ws_bytes = int(lib.rdg_densify_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    src_map = torch.empty(3 * n, dtype=torch.int32, device=dev)
    split_rank = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    check(lib.rdg_densify_plan(n, ptr(scaling), sw, ptr(opacity), ptr(stats.grad_accum), ptr(stats.denom),
                               float(grad_threshold), float(percent_dense), float(extent), float(min_opacity),
                               1 if max_screen_size else 0, ptr(src_map), ptr(split_rank), ptr(counts), ptr(ws), ws_bytes, stream))
    n_a, n_b, n_c, n_s = (int(v) for v in counts.tolist())            # the one host read (new sizes)
    n_new = n_a + n_b + 2 * n_c
    if noise is None:
        noise = torch.randn(max(2 * n_s, 1), 3, device=dev, generator=generator)
    elif n_s > 0 and (noise.shape[0] != 2 * n_s or not noise.is_cuda or not noise.is_contiguous()):
        raise RuntimeError(f"noise must be a contiguous CUDA tensor [2 * {n_s}, 3]")

    fields = []
    new_params, new_moments, new_extras = {}, ({} if moments is not None else None), {}

    def add_field(src, mode):
        dst = torch.empty((n_new,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
        fields.append((src, dst, _rows(src)[1], mode))
        return dst

    for name, t in params.items():
        new_params[name] = add_field(t, PARAM_MODES.get(name, MODE_COPY))
        if moments is not None and name in moments:
            m, v = moments[name]
            new_moments[name] = (add_field(m, MODE_MOMENT), add_field(v, MODE_MOMENT))
    for name, t in (extras or {}).items():
        new_extras[name] = add_field(t, MODE_COPY)
    # at most 32 fields per launch
    for lo in range(0, len(fields) if n_new else 0, 32):
        chunk = fields[lo:lo + 32]
        arr = (_lib.RdgDensifyField * len(chunk))()
        for k, (src, dst, w, mode) in enumerate(chunk):
            arr[k].src, arr[k].dst, arr[k].width, arr[k].mode = ptr(src), ptr(dst), w, mode
        check(lib.rdg_densify_apply(n_new, ptr(src_map), ptr(split_rank), ptr(counts), ptr(noise), arr, len(chunk), ptr(xyz),
                                    ptr(scaling), sw, ptr(rotation), stream))
    info = {"survivors": n_a, "clones": n_b, "split_kept": n_c, "split_selected": n_s, "rows": n_new}
    return new_params, new_moments, new_extras, DensifyStats(n_new, dev), info


def reset_opacity(opacity: torch.Tensor, moments: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, cap: float = 0.01):
    """ThreeDGSTrainer.reset_opacity (rodygs_static.py:150-159), in place; the group's moments are zeroed."""
    _lib.require_cuda(opacity)
    m, v = moments if moments is not None else (None, None)
    check(_lib.load().rdg_reset_opacity(opacity.numel(), ptr(opacity), float(cap), ptr(m), ptr(v), _lib.stream_ptr()))
    return opacity


def should_densify(iteration: int, densify_from_iter: int, densify_until_iter: int, densification_interval: int) -> bool:
    """The gate of rodygs.py:317,343-347."""
    return (iteration < densify_until_iter and densification_interval != 0 and iteration > densify_from_iter
            and iteration % densification_interval == 0)


def size_threshold(iteration: int, opacity_reset_interval: int) -> Optional[int]:
    """rodygs.py:348-350."""
    return 20 if iteration > opacity_reset_interval else None
