"""Adam on the flat parameter buffer with the reference's parameter groups.

Mirror of ThreeDGSTrainer.optim_setup / update_learning_rate
(/root/reference/src/trainer/rodygs_static.py:106-149), DynTrainer.append_motion_optim
(/root/reference/src/trainer/rodygs_dynamic.py:92-123) and get_expon_lr_func
(/root/reference/src/utils/general_utils.py:40-73).  One CUDA launch per optimiser step
(rdg_adam_groups); the deformation MLP (68 K parameters) stays with torch.optim.Adam.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr

# flat-layout field -> the reference's group name
GROUP_OF = {"xyz": "xyz", "features_dc": "f_dc", "features_rest": "f_rest", "opacity": "opacity", "scaling": "scaling",
            "rotation": "rotation"}


def expon_lr(step: int, lr_init: float, lr_final: float, lr_delay_steps: int = 0, lr_delay_mult: float = 1.0,
             max_steps: int = 1000000) -> float:
    """get_expon_lr_func(...)(step) (general_utils.py:40-73): log-linear decay with an optional eased-in start."""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0.0), 1.0)
    return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


@dataclass
class GaussianLRs:
    """train_kubric_mrig.yaml:159-167 (static) / :236-260 (dynamic: scaling_lr 0.001, motion_coeff_lr 1.6e-4)."""
    position_lr_init: float = 0.00016
    position_lr_final: float = 1.6e-06
    position_lr_delay_mult: float = 0.01
    position_lr_max_steps: int = 20000
    feature_lr: float = 0.0025
    opacity_lr: float = 0.05
    scaling_lr: float = 0.005
    rotation_lr: float = 0.001
    motion_coeff_lr: Optional[float] = None      # dynamic model only

    def group_lrs(self, iteration: int, spatial_lr_scale: float) -> Dict[str, float]:
        """Learning rate of every group at `iteration` (only xyz is scheduled, rodygs_static.py:143-149)."""
        lrs = {
            "xyz": expon_lr(iteration, self.position_lr_init * spatial_lr_scale, self.position_lr_final * spatial_lr_scale,
                            lr_delay_mult=self.position_lr_delay_mult, max_steps=self.position_lr_max_steps),
            "f_dc": self.feature_lr, "f_rest": self.feature_lr / 20.0, "opacity": self.opacity_lr,
            "scaling": self.scaling_lr, "rotation": self.rotation_lr,
        }
        if self.motion_coeff_lr is not None:
            lrs["motion_coeff"] = self.motion_coeff_lr
        return lrs


class GaussianAdam:
    """torch.optim.Adam(groups, lr=0.0, eps=1e-15) of ONE model ("static" or "dynamic") living in the flat buffers
    of a SplatTrainStep.  `ranges`: reference group name -> (float offset, numel) inside those buffers."""

    def __init__(self, ranges: Dict[str, Tuple[int, int]], total: int, lrs: GaussianLRs, spatial_lr_scale: float,
                 device="cuda", betas=(0.9, 0.999), eps: float = 1e-15):
        self.ranges = dict(ranges)
        self.lrs, self.spatial_lr_scale = lrs, float(spatial_lr_scale)
        self.betas, self.eps = betas, eps
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=device)
        self.steps = 0

    def moments(self, name: str) -> Tuple[torch.Tensor, torch.Tensor]:
        off, n = self.ranges[name]
        return self.exp_avg[off:off + n], self.exp_avg_sq[off:off + n]

    def step(self, params: torch.Tensor, grads: torch.Tensor, iteration: int, grad_scale: float = 1.0, skip=()):
        """update_learning_rate(iteration) + optimizer.step() (rodygs.py:209,364).  skip: group names another kernel has
        already stepped for this iteration (the SH groups, SplatTrainStep._sh_adam_from_factors)."""
        _lib.require_cuda(params, grads)
        if params.numel() != self.exp_avg.numel() or grads.numel() != params.numel():
            raise RuntimeError("optimizer state does not match the flat buffers (rebuild it after densification)")
        lr = self.lrs.group_lrs(iteration, self.spatial_lr_scale)
        live = [(name, off, n) for name, (off, n) in self.ranges.items() if n > 0 and name not in skip]
        arr = (_lib.RdgAdamGroup * len(live))()
        for k, (name, off, n) in enumerate(live):
            arr[k].begin, arr[k].end, arr[k].lr = off, off + n, lr[name]
        self.steps += 1
        if live:
            check(_lib.load().rdg_adam_groups(ptr(params), ptr(grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), arr, len(live),
                                              self.betas[0], self.betas[1], self.eps, self.steps, float(grad_scale),
                                              _lib.stream_ptr()))
        return lr
