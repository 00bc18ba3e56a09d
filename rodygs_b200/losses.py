"""Fused photometric and depth losses, same names and argument meaning as
/root/reference/src/utils/loss_utils.py (`l1_loss` :19-20, `ssim` :57-66,
`pearson_depth_loss` :100-117) and the weighted sum that
/root/reference/src/trainer/losses.py:61-107 builds from
configs/train/train_kubric_mrig.yaml:135-144 (0.2 * (1 - ssim) + 0.8 * l1).

One CUDA call computes the loss *and* its gradient w.r.t. the prediction
(rodygs_b200/csrc/loss.cu); autograd only scales that gradient.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import check, ptr


def _img3(t: torch.Tensor) -> torch.Tensor:
    if t.dim() == 4:
        if t.shape[0] != 1:
            raise Exception("batched images are not supported; pass [C,H,W]")
        t = t[0]
    if t.dim() != 3:
        raise Exception("expected an image of shape [C,H,W]")
    if not t.is_cuda:
        raise RuntimeError("rodygs_b200 runs on CUDA tensors only (no CPU fallback)")
    return t.float().contiguous()


class _PhotometricFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, w_l1, w_dssim, need_grad):
        lib = _lib.load()
        pred, gt = _img3(pred), _img3(gt)
        if pred.shape != gt.shape:
            raise Exception("pred and gt must have the same shape")
        ch, h, w = pred.shape
        ws_bytes = int(lib.rdg_l1_dssim_workspace_bytes(ch, h, w))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pred.device)
        out = torch.zeros(3, dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred) if need_grad else None
        check(lib.rdg_l1_dssim(ptr(pred), ptr(gt), ch, h, w, float(w_l1), float(w_dssim), ptr(out), ptr(grad),
                               ptr(ws), ws_bytes, _lib.stream_ptr()))
        ctx.grad = grad
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        if ctx.grad is None:
            raise RuntimeError("photometric loss was computed without a gradient")
        return ctx.grad * g_loss, None, None, None, None


def photometric_loss(pred: torch.Tensor, gt: torch.Tensor, w_l1: float = 0.8, w_dssim: float = 0.2,
                     return_parts: bool = False):
    """w_l1 * mean|pred-gt| + w_dssim * (1 - SSIM(pred, gt)); parts = (loss, l1, ssim)."""
    loss, parts = _PhotometricFn.apply(pred, gt, w_l1, w_dssim, pred.requires_grad)
    return (loss, parts) if return_parts else loss


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    return photometric_loss(network_output, gt, 1.0, 0.0)


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    if window_size != 11 or not size_average:
        raise NotImplementedError("only window_size=11, size_average=True (the reference's only use) is implemented")
    # 1 - ((w_dssim=1) * (1 - ssim)) == ssim, differentiable through the fused kernel
    return 1.0 - photometric_loss(img1, img2, 0.0, 1.0)


class _PearsonFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, boxes, weights, eps, need_grad):
        lib = _lib.load()
        p = pred.float().contiguous()
        g = gt.float().contiguous()
        if not p.is_cuda:
            raise RuntimeError("rodygs_b200 runs on CUDA tensors only (no CPU fallback)")
        h, w = p.shape[-2], p.shape[-1]
        nb = boxes.shape[0]
        out = torch.empty(nb, dtype=torch.float32, device=p.device)
        stats = torch.empty(nb * 8, dtype=torch.float64, device=p.device)
        grad = torch.zeros_like(p) if need_grad else None
        check(lib.rdg_pearson(ptr(p), ptr(g), h, w, ptr(boxes), ptr(weights), nb, float(eps), ptr(out), ptr(grad),
                              ptr(stats), _lib.stream_ptr()))
        ctx.grad = grad
        ctx.pred_shape = pred.shape
        loss = (out * weights).sum() if weights is not None else out.sum()
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        return (ctx.grad * g_loss).reshape(ctx.pred_shape), None, None, None, None, None


def pearson_depth_loss(input_depth: torch.Tensor, target_depth: torch.Tensor, eps: float = 1e-6, mask=None):
    """GlobalPearsonDepthLoss body (losses.py:110-129, mode 'all' => mask None)."""
    if mask is not None:
        raise NotImplementedError("masked Pearson is not used by the reference configs (mode: all)")
    h, w = input_depth.shape[-2], input_depth.shape[-1]
    boxes = torch.tensor([[0, 0, h, w]], dtype=torch.int32, device=input_depth.device)
    return _PearsonFn.apply(input_depth, target_depth, boxes, None, eps, input_depth.requires_grad)


def local_pearson_depth_loss(pred_depth: torch.Tensor, gt_depth: torch.Tensor, box_p: int = 128, p_corr: float = 0.5,
                             eps: float = 1e-6, generator: Optional[torch.Generator] = None,
                             origins: Optional[torch.Tensor] = None):
    """LocalPearsonDepthLoss (losses.py:132-182) as ONE batched launch over all boxes
    instead of a Python loop; box origins are drawn on the device like the reference
    (`torch.randint(0, max_h, ...)`, `torch.randint(0, max_w, ...)`)."""
    h, w = pred_depth.shape[-2], pred_depth.shape[-1]
    dev = pred_depth.device
    if origins is None:
        n_corr = int(p_corr * math.floor(h / box_p) * math.floor(w / box_p))
        if n_corr == 0:
            return pred_depth.new_zeros(())
        x0 = torch.randint(0, h - box_p, (n_corr,), device=dev, generator=generator)
        y0 = torch.randint(0, w - box_p, (n_corr,), device=dev, generator=generator)
    else:
        x0, y0 = origins[:, 0].to(dev), origins[:, 1].to(dev)
        n_corr = x0.shape[0]
    boxes = torch.stack([x0, y0, torch.full_like(x0, box_p), torch.full_like(x0, box_p)], 1).to(torch.int32).contiguous()
    weights = torch.full((n_corr,), 1.0 / n_corr, dtype=torch.float32, device=dev)
    return _PearsonFn.apply(pred_depth, gt_depth, boxes, weights, eps, pred_depth.requires_grad)
