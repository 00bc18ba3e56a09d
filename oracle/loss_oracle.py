"""PyTorch-CPU restatement of the photometric / depth losses on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the
reference's own Python through ``tests/golden/losses_*.npz``.

Follows
* ``/root/reference/src/utils/loss_utils.py:19-20``   (l1_loss)
* ``/root/reference/src/utils/loss_utils.py:34-97``   (gaussian / create_window / ssim / _ssim)
* ``/root/reference/src/utils/loss_utils.py:100-117`` (pearson_depth_loss)
* ``/root/reference/src/trainer/losses.py:61-107``    (MultiLoss weights: 0.2*(1-ssim) + 0.8*l1,
  ``configs/train/train_kubric_mrig.yaml:135-144``)
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

WINDOW = 11
SIGMA = 1.5
SSIM_C1 = 0.01 ** 2
SSIM_C2 = 0.03 ** 2


def gauss_taps(dtype=torch.float32) -> torch.Tensor:
    """11 normalised taps, sigma 1.5 (loss_utils.py:34-41).  The reference builds
    them in float32 (torch.Tensor of Python floats, then / sum)."""
    g = torch.tensor([math.exp(-((i - WINDOW // 2) ** 2) / float(2 * SIGMA ** 2))
                      for i in range(WINDOW)], dtype=torch.float32)
    return (g / g.sum()).to(dtype)


def _blur(img: torch.Tensor, w2d: torch.Tensor) -> torch.Tensor:
    ch = img.shape[-3]
    return F.conv2d(img, w2d.expand(ch, 1, WINDOW, WINDOW), padding=WINDOW // 2, groups=ch)


def ssim_map(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """[C,H,W] (or [B,C,H,W]) -> per-pixel SSIM, zero padding 5, per channel
    (loss_utils.py:68-92)."""
    t = gauss_taps(img1.dtype)
    w2d = torch.outer(t, t)[None, None]
    a, b = (img1, img2) if img1.dim() == 4 else (img1[None], img2[None])
    mu1, mu2 = _blur(a, w2d), _blur(b, w2d)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _blur(a * a, w2d) - mu1_sq
    s2 = _blur(b * b, w2d) - mu2_sq
    s12 = _blur(a * b, w2d) - mu12
    m = ((2 * mu12 + SSIM_C1) * (2 * s12 + SSIM_C2)) / ((mu1_sq + mu2_sq + SSIM_C1) * (s1 + s2 + SSIM_C2))
    return m if img1.dim() == 4 else m[0]


def ssim(img1, img2):
    return ssim_map(img1, img2).mean()


def l1(img1, img2):
    return (img1 - img2).abs().mean()


def photometric(pred, gt, w_l1=0.8, w_dssim=0.2):
    """0.8*L1 + 0.2*(1-SSIM) (losses.py:61-107 with the yaml weights)."""
    return w_l1 * l1(pred, gt) + w_dssim * (1.0 - ssim(pred, gt))


def pearson_depth(pred, gt, eps=1e-6):
    """1 - mean(z_pred*z_gt), z = (d-mean)/(std_unbiased+eps) (loss_utils.py:100-117)."""
    p = pred.reshape(-1)
    g = gt.reshape(-1)
    pc = p - p.mean()
    gc = g - g.mean()
    pn = pc / (pc.std() + eps)
    gn = gc / (gc.std() + eps)
    return 1.0 - (pn * gn).mean()


def local_pearson_depth(pred, gt, x0, y0, box_p=128, eps=1e-6):
    """Mean over the given boxes of pearson_depth (losses.py:132-182); the box
    origins (x0 = row, y0 = col, as named in the reference) are inputs here so
    that both sides use the same random draw."""
    tot = pred.new_zeros(())
    for r, c in zip(x0.tolist(), y0.tolist()):
        tot = tot + pearson_depth(pred[:, r:r + box_p, c:c + box_p], gt[:, r:r + box_p, c:c + box_p], eps)
    return tot / max(len(x0), 1)
