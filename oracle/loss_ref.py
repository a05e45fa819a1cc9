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
