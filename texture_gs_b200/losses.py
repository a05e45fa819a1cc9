"""SURVEY §8f N3 — fused photometric loss of the reference's training step.

``photometric_loss(image, gt, lambda_dssim)`` returns ``(loss, Ll1, Lssim)`` exactly as
``models/texture_gaussian3d.py:333-340`` combines ``losses.l1_loss`` (``losses/pixelwise_loss.py:3-4``)
and ``1 - losses.ssim_loss`` (``losses/ssim_loss.py:16-54``):

    loss = (1 - lambda_dssim) * Ll1 + lambda_dssim * Lssim

in two CUDA kernels forward (+1 tiny finalize) and one backward, instead of the reference's five
depthwise 11x11 convolutions, ~10 pointwise kernels and their autograd twins. No CPU fallback.

``geometry_losses(alpha, norm, gt_alpha, gt_norm, gt_image, gamma)`` returns ``(Lalpha, Lnorm, Lnsm)`` of the
same step (``models/texture_gaussian3d.py:342-345, 354-358, 365-368``: ``l1_loss``, ``norm_loss``
``losses/norm_reg_loss.py:66-71``, ``smooth_loss`` ``losses/smooth_loss.py:4-27``) from one kernel forward
(+ finalize) and one backward, differentiable w.r.t. the rasterizer outputs ``alpha`` and ``norm``.
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


def _opt(t, dev):
    return None if t is None else t.detach().to(dev).float().contiguous()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _GeometryLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, norm, gt_alpha, gt_norm, gt_image, gamma: float):
        lib = L.load()
        if not alpha.is_cuda or not norm.is_cuda:
            raise L.TexgsError("geometry_losses runs on CUDA tensors only (no CPU fallback)")
        if alpha.dim() != 3 or alpha.shape[0] != 1 or norm.dim() != 3 or norm.shape[0] != 3 or alpha.shape[1:] != norm.shape[1:]:
            raise L.TexgsError(f"alpha must be (1,H,W) and norm (3,H,W); got {tuple(alpha.shape)} and {tuple(norm.shape)}")
        dev = alpha.device
        H, W = alpha.shape[1:]
        a, n = _opt(alpha, dev), _opt(norm, dev)
        ga, gn, gi = _opt(gt_alpha, dev), _opt(gt_norm, dev), _opt(gt_image, dev)
        for name, t, ch in (("gt_alpha", ga, 1), ("gt_norm", gn, 3), ("gt_image", gi, 3)):
            if t is not None and tuple(t.shape) != (ch, H, W):
                raise L.TexgsError(f"{name} must be ({ch},{H},{W}), got {tuple(t.shape)}")
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            nbytes = C.c_size_t()
            L.check(lib.texgs_geometry_loss_workspace_size(H, W, C.byref(nbytes)), "texgs_geometry_loss_workspace_size")
            ws = torch.empty(nbytes.value, device=dev, dtype=torch.uint8)
            out3 = torch.empty(3, device=dev, dtype=torch.float32)
            L.check(lib.texgs_geometry_loss_forward(_p(a), _p(n), _p(ga), _p(gn), _p(gi), H, W, float(gamma), _p(ws), _p(out3),
                                                    C.c_void_p(stream)), "texgs_geometry_loss_forward")
        ctx.tensors = (a, n, ga, gn, gi, ws)
        ctx.gamma = float(gamma)
        return out3[0], out3[1], out3[2]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_alpha, g_norm, g_nsm):
        lib = L.load()
        a, n, ga, gn, gi, ws = ctx.tensors
        dev = a.device
        H, W = a.shape[1:]
        z = torch.zeros((), device=dev)
        coef = torch.stack([z if g is None else g.float() for g in (g_alpha, g_norm, g_nsm)]).contiguous()
        need_a, need_n = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_a = torch.empty_like(a) if need_a else None
        d_n = torch.empty_like(n) if need_n else None
        if need_a or need_n:
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                L.check(lib.texgs_geometry_loss_backward(_p(a), _p(n), _p(ga), _p(gn), _p(gi), H, W, ctx.gamma, _p(ws), _p(coef),
                                                         _p(d_a), _p(d_n), C.c_void_p(stream)), "texgs_geometry_loss_backward")
        return d_a, d_n, None, None, None, None


def geometry_losses(alpha, norm, gt_alpha=None, gt_norm=None, gt_image=None, gamma: float = 0.1):
    """(Lalpha, Lnorm, Lnsm) of ``models/texture_gaussian3d.py:342-368``; ``gt_alpha=None`` means ones
    (``:330``), ``gt_norm`` / ``gt_image`` ``None`` skips that loss (it reads 0). Differentiable w.r.t. ``alpha``
    and ``norm``."""
    return _GeometryLosses.apply(alpha, norm, gt_alpha, gt_norm, gt_image, float(gamma))


_CONSTS: dict = {}


def _const_vector(dev, values):
    """small per-device constant vectors (loss weights), uploaded once"""
    key = (str(dev), tuple(float(v) for v in values))
    t = _CONSTS.get(key)
    if t is None:
        t = _CONSTS[key] = torch.tensor(key[1], dtype=torch.float32, device=dev)
    return t


class _TrainingLoss(torch.autograd.Function):
    """loss = (1-l) L1 + l (1-SSIM) + l_alpha Lalpha + l_norm Lnorm + l_nsm Lnsm in one autograd node: the two fused forward
    kernels, ONE small torch op to weight the six partial losses, and in the backward one op to scale the incoming gradient
    into the coefficient vectors of the two fused backward kernels — instead of ~20 scalar glue kernels per view around
    ``photometric_loss`` + ``geometry_losses`` + the weighted sum of models/texture_gaussian3d.py:333-368."""

    @staticmethod
    def forward(ctx, image, alpha, norm, gt_image, gt_alpha, gt_norm, lam, l_alpha, l_norm, l_nsm, gamma):
        lib = L.load()
        if not (image.is_cuda and alpha.is_cuda and norm.is_cuda):
            raise L.TexgsError("training_loss runs on CUDA tensors only (no CPU fallback)")
        dev = image.device
        img, a, n = _opt(image, dev), _opt(alpha, dev), _opt(norm, dev)
        gi, ga, gn = _opt(gt_image, dev), _opt(gt_alpha, dev), _opt(gt_norm, dev)
        Cc, H, W = img.shape
        if tuple(a.shape) != (1, H, W) or tuple(n.shape) != (3, H, W) or tuple(gi.shape) != (Cc, H, W):
            raise L.TexgsError("training_loss: image / gt_image (3,H,W), alpha (1,H,W), norm (3,H,W) expected")
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            nb = C.c_size_t()
            L.check(lib.texgs_photometric_workspace_size(Cc, H, W, C.byref(nb)), "texgs_photometric_workspace_size")
            ws_p = torch.empty(nb.value, device=dev, dtype=torch.uint8)
            L.check(lib.texgs_geometry_loss_workspace_size(H, W, C.byref(nb)), "texgs_geometry_loss_workspace_size")
            ws_g = torch.empty(nb.value, device=dev, dtype=torch.uint8)
            out6 = torch.empty(6, device=dev, dtype=torch.float32)      # [combined photometric, L1, 1-SSIM, Lalpha, Lnorm, Lnsm]
            L.check(lib.texgs_photometric_forward(_p(img), _p(gi), Cc, H, W, float(lam), _p(ws_p), _p(out6), st), "texgs_photometric_forward")
            L.check(lib.texgs_geometry_loss_forward(_p(a), _p(n), _p(ga), _p(gn), _p(gi), H, W, float(gamma), _p(ws_g), _p(out6[3:]), st),
                    "texgs_geometry_loss_forward")
        w6 = _const_vector(dev, (1.0, 0.0, 0.0, l_alpha, l_norm, l_nsm))
        ctx.tensors = (img, a, n, gi, ga, gn, ws_p, ws_g)
        ctx.consts = (float(lam), float(l_alpha), float(l_norm), float(l_nsm), float(gamma))
        ctx.mark_non_differentiable(out6)
        return torch.dot(out6, w6), out6

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss, _g_parts):
        lib = L.load()
        img, a, n, gi, ga, gn, ws_p, ws_g = ctx.tensors
        lam, l_alpha, l_norm, l_nsm, gamma = ctx.consts
        dev = img.device
        Cc, H, W = img.shape
        coef = g_loss.float().reshape(1) * _const_vector(dev, (1.0 - lam, lam, l_alpha, l_norm, l_nsm))
        need = ctx.needs_input_grad
        d_img = torch.empty_like(img) if need[0] else None
        d_a = torch.empty_like(a) if need[1] else None
        d_n = torch.empty_like(n) if need[2] else None
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if d_img is not None:
                L.check(lib.texgs_photometric_backward(_p(img), _p(gi), Cc, H, W, _p(ws_p), _p(coef), _p(d_img), st), "texgs_photometric_backward")
            if d_a is not None or d_n is not None:
                L.check(lib.texgs_geometry_loss_backward(_p(a), _p(n), _p(ga), _p(gn), _p(gi), H, W, gamma, _p(ws_g), _p(coef[2:]), _p(d_a), _p(d_n), st),
                        "texgs_geometry_loss_backward")
        return d_img, d_a, d_n, None, None, None, None, None, None, None, None


def training_loss(image, alpha, norm, gt_image, gt_alpha, gt_norm, lambda_dssim: float = 0.2, lambda_alpha: float = 1.0,
                  lambda_norm: float = 0.1, lambda_norm_smooth: float = 0.5, gamma: float = 0.1):
    """The image-space part of ``compute_loss`` (``models/texture_gaussian3d.py:333-368`` with the losses
    ``configs/texture_gaussian3d.yaml:77-88`` enables) as ONE differentiable scalar plus its six parts
    ``[photometric, Ll1, Lssim, Lalpha, Lnorm, Lnsm]`` (detached, for ``loss_stats``):
    ``(1-l) Ll1 + l Lssim + lambda_alpha Lalpha + lambda_norm Lnorm + lambda_norm_smooth Lnsm``."""
    return _TrainingLoss.apply(image, alpha, norm, gt_image, gt_alpha, gt_norm, float(lambda_dssim), float(lambda_alpha),
                               float(lambda_norm), float(lambda_norm_smooth), float(gamma))
