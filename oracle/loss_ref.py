"""ORACLE (test infrastructure, NOT product code) — restatement of the reference's photometric loss.

PINNED: unlike the rasterizer oracle, this one restates code that IS in the reference tree and is
checked against golden vectors produced by that code itself (tests/golden/photometric_loss.npz,
generator tests/golden/make_loss_golden.py imports /root/reference/losses directly).

  l1_loss     — reference losses/pixelwise_loss.py:3-4     mean |x - y|
  ssim_loss   — reference losses/ssim_loss.py:6-54         11x11 Gaussian window (sigma 1.5, separable
                weights normalised in fp32), zero padding 5, per-channel (groups), C1=0.01^2, C2=0.03^2,
                ssim_map.mean()
  combination — reference models/texture_gaussian3d.py:333-340
                loss = (1 - lambda_dssim) * L1 + lambda_dssim * (1 - SSIM)
  norm_loss   — reference losses/norm_reg_loss.py:66-71    masked mean of 1 - <pred, gt>
  smooth_loss — reference losses/smooth_loss.py:4-27       bilateral first-order smoothness, 4 directions
  geometry_losses — the (Lalpha, Lnorm, Lnsm) triple of models/texture_gaussian3d.py:342-368
                (golden vectors: tests/golden/geometry_loss.npz, generator make_geometry_loss_golden.py)
"""
from __future__ import annotations

from math import exp

import torch
import torch.nn.functional as F

WINDOW = 11
SIGMA = 1.5
C1 = 0.01 ** 2
C2 = 0.03 ** 2


def gaussian_window_1d(dtype=torch.float32) -> torch.Tensor:
    g = torch.tensor([exp(-(x - WINDOW // 2) ** 2 / float(2 * SIGMA ** 2)) for x in range(WINDOW)], dtype=torch.float32)
    return (g / g.sum()).to(dtype)


def l1_loss(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return (x - y).abs().mean()


def ssim(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x, y: (C,H,W). Mean SSIM exactly as the reference computes it (2-D window = outer product of
    the fp32 1-D window, zero padding)."""
    C = x.shape[-3]
    w1 = gaussian_window_1d(torch.float32)
    w2 = (w1[:, None] @ w1[None, :]).to(dtype=x.dtype, device=x.device)
    win = w2.expand(C, 1, WINDOW, WINDOW).contiguous()
    a, b = x[None], y[None]
    conv = lambda t: F.conv2d(t, win, padding=WINDOW // 2, groups=C)
    mu1, mu2 = conv(a), conv(b)
    s11 = conv(a * a) - mu1 * mu1
    s22 = conv(b * b) - mu2 * mu2
    s12 = conv(a * b) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))
    return m.mean()


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float):
    """Returns (loss, Ll1, Lssim) with Lssim = 1 - SSIM, as models/texture_gaussian3d.py:333-340."""
    ll1 = l1_loss(image, gt)
    lssim = 1.0 - ssim(image, gt)
    return (1.0 - lambda_dssim) * ll1 + lambda_dssim * lssim, ll1, lssim


def norm_loss(pred: torch.Tensor, gt: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
    """(3,H,W) normals; with a mask: sum((1 - <pred,gt>) * mask) / (sum(mask) + 1e-6)."""
    cos = (pred * gt).sum(dim=0, keepdim=True)
    if mask is None:
        return (1.0 - cos).mean()
    return ((1.0 - cos) * mask).sum() / (mask.sum() + 1e-6)


def smooth_loss(rgb: torch.Tensor, value: torch.Tensor, mask: torch.Tensor = None, gamma: float = 0.1) -> torch.Tensor:
    """Bilateral smoothness of ``value`` (C,H,W) guided by ``rgb`` (3,H,W): for the neighbour directions
    right, down, down-right and up-right, weight = exp(-sum_c|rgb_a - rgb_b| / gamma) * mask_a * mask_b and
    L_d = sum|w * (value_a - value_b)| / (sum(w) + 1e-6); returns the mean of the four."""
    H, W = rgb.shape[-2:]
    views = [  # (a, b) index pairs of each direction
        ((slice(None), slice(0, W - 1)), (slice(None), slice(1, W))),
        ((slice(0, H - 1), slice(None)), (slice(1, H), slice(None))),
        ((slice(0, H - 1), slice(0, W - 1)), (slice(1, H), slice(1, W))),
        ((slice(1, H), slice(0, W - 1)), (slice(0, H - 1), slice(1, W))),
    ]
    total = 0.0
    for (ay, ax), (by, bx) in views:
        w = torch.exp(-(rgb[:, ay, ax] - rgb[:, by, bx]).abs().sum(0, keepdim=True) / gamma)
        if mask is not None:
            m = mask.to(w.dtype)
            w = w * (m[:, ay, ax] * m[:, by, bx])
        total = total + (w * (value[:, ay, ax] - value[:, by, bx])).abs().sum() / (w.sum() + 1e-6)
    return total / 4


def geometry_losses(alpha, norm, gt_alpha, gt_norm, gt_image, gamma: float = 0.1):
    """(Lalpha, Lnorm, Lnsm) as models/texture_gaussian3d.py:342-345, 354-358, 365-368 compute them."""
    return l1_loss(alpha, gt_alpha), norm_loss(norm, gt_norm, gt_alpha), smooth_loss(gt_image, norm, gt_alpha, gamma)
