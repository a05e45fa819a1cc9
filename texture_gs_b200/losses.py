"""SURVEY §8f N3 — fused photometric loss of the reference's training step.

``photometric_loss(image, gt, lambda_dssim)`` returns ``(loss, Ll1, Lssim)`` exactly as
``models/texture_gaussian3d.py:333-340`` combines ``losses.l1_loss`` (``losses/pixelwise_loss.py:3-4``)
and ``1 - losses.ssim_loss`` (``losses/ssim_loss.py:16-54``):

    loss = (1 - lambda_dssim) * Ll1 + lambda_dssim * Lssim

in two CUDA kernels forward (+1 tiny finalize) and one backward, instead of the reference's five
depthwise 11x11 convolutions, ~10 pointwise kernels and their autograd twins. No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float):
        lib = L.load()
        if not image.is_cuda:
            raise L.TexgsError("photometric_loss runs on CUDA tensors only (no CPU fallback)")
        if image.shape != gt.shape or image.dim() != 3:
            raise L.TexgsError(f"image and gt must both be (C,H,W); got {tuple(image.shape)} and {tuple(gt.shape)}")
        dev = image.device
        img = image.detach().float().contiguous()
        ref = gt.detach().to(dev).float().contiguous()
        Cc, H, W = img.shape
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            nbytes = C.c_size_t()
            L.check(lib.texgs_photometric_workspace_size(Cc, H, W, C.byref(nbytes)), "texgs_photometric_workspace_size")
            ws = torch.empty(nbytes.value, device=dev, dtype=torch.uint8)
            out3 = torch.empty(3, device=dev, dtype=torch.float32)
            L.check(lib.texgs_photometric_forward(C.c_void_p(img.data_ptr()), C.c_void_p(ref.data_ptr()), Cc, H, W,
                                                  float(lambda_dssim), C.c_void_p(ws.data_ptr()), C.c_void_p(out3.data_ptr()),
                                                  C.c_void_p(stream)), "texgs_photometric_forward")
        ctx.save_for_backward(img, ref, ws)
        ctx.lam = float(lambda_dssim)
        return out3[0], out3[1], out3[2]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss, g_l1, g_lssim):
        lib = L.load()
        img, ref, ws = ctx.saved_tensors
        dev = img.device
        Cc, H, W = img.shape
        z = torch.zeros((), device=dev)
        g_loss = z if g_loss is None else g_loss.float()
        g_l1 = z if g_l1 is None else g_l1.float()
        g_lssim = z if g_lssim is None else g_lssim.float()
        coef = torch.stack([g_loss * (1.0 - ctx.lam) + g_l1, g_loss * ctx.lam + g_lssim]).contiguous()
        dimg = torch.empty_like(img)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            L.check(lib.texgs_photometric_backward(C.c_void_p(img.data_ptr()), C.c_void_p(ref.data_ptr()), Cc, H, W,
                                                   C.c_void_p(ws.data_ptr()), C.c_void_p(coef.data_ptr()),
                                                   C.c_void_p(dimg.data_ptr()), C.c_void_p(stream)), "texgs_photometric_backward")
        return dimg, None, None


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float):
    """(loss, Ll1, Lssim) of ``models/texture_gaussian3d.py:333-340``; differentiable w.r.t. ``image``."""
    return _PhotometricLoss.apply(image, gt, float(lambda_dssim))
