"""SURVEY §8f N1 — the UV + Jacobian producer of the training step, one kernel.

The reference evaluates ``uv = uv_net(xyz, geo_emb)`` (``TextureGaussian3D.get_uvs``,
``models/texture_gaussian3d.py:230-236``; network ``models/modules/uv_net.py:8-36``) and then
``grad_uvs = jacobian(lambda x: uv_net(x, emb).sum(0), xyz)`` (``get_grad_uvs``, ``:217-227``): one forward and three
backward passes through the MLP every iteration. ``FusedUVNet.uv_and_jacobian`` returns both from a single tcgen05
kernel (value + three forward-mode tangents per point are four rows of the same GEMMs; fp16 operands, fp32
accumulation — the reference's tiny-cuda-nn path computes in fp16 as well).

``FusedUVNet`` keeps its parameters in the layout of the reference's ``nn.Linear`` networks
(``models/modules/utils.py:44-55``; ``state_dict`` keys ``pre_mlp.{0,2}``, ``mlp.{0,2,4}``), with ``bias=False``
for tiny-cuda-nn-style networks. ``uv`` is differentiable w.r.t. ``xyz``, ``emb`` and all parameters (backward =
plain fp16 GEMMs on the activations the kernel stashed); ``grad_uvs`` carries no gradient, as in the reference
(``:227``). No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import nn

from . import _lib as L

HIDDEN = 128


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t, dev):
    return None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()


class _UvMlp(torch.autograd.Function):
    """inputs: xyz, emb, W1, b1, W2, b2, W3, b3, W4, b4, W5, b5 (biases may be None), offset, inv_scale (3 floats each)."""

    @staticmethod
    def forward(ctx, xyz, emb, W1, b1, W2, b2, W3, b3, W4, b4, W5, b5, offset, inv_scale, debug):
        lib = L.load()
        if not xyz.is_cuda:
            raise L.TexgsError("FusedUVNet runs on CUDA tensors only (no CPU fallback); got " + str(xyz.device))
        if xyz.dim() != 2 or xyz.shape[1] != 3:
            raise L.TexgsError(f"xyz must be (N,3), got {tuple(xyz.shape)}")
        for name, w, shape in (("W1", W1, (HIDDEN, 3)), ("W2", W2, (HIDDEN, HIDDEN)), ("W3", W3, (HIDDEN, HIDDEN)),
                               ("W4", W4, (HIDDEN, HIDDEN)), ("W5", W5, (3, HIDDEN)), ("emb", emb, (HIDDEN,))):
            if tuple(w.shape) != shape:
                raise L.TexgsError(f"{name} must be {shape}, got {tuple(w.shape)} (the kernel is built for the 128-wide network "
                                   "of configs/texture_gaussian3d.yaml:18-27)")
        dev = xyz.device
        N = xyz.shape[0]
        x = _f32(xyz, dev)
        w1, e = _f32(W1, dev), _f32(emb, dev)
        wh = [w.detach().to(device=dev, dtype=torch.float16).contiguous() for w in (W2, W3, W4, W5)]
        bs = [_f32(b, dev) for b in (b1, b2, b3, b4, b5)]
        need_grad = any(ctx.needs_input_grad[:12])
        uv = torch.empty(N, 3, device=dev, dtype=torch.float32)
        jac = torch.empty(N, 9, device=dev, dtype=torch.float32)
        stash = [torch.empty(N, HIDDEN, device=dev, dtype=torch.float16) for _ in range(4)] if need_grad else [None] * 4
        inv_len = torch.empty(N, device=dev, dtype=torch.float32) if need_grad else None
        dbg = torch.zeros(4 * 128 * 128 + 128 * 16, device=dev, dtype=torch.float32) if debug else None
        a = L.TexgsUvMlpArgs()
        a.N = N
        a.xyz = _p(x)
        a.offset = (C.c_float * 3)(*[float(v) for v in offset])
        a.inv_scale = (C.c_float * 3)(*[float(v) for v in inv_scale])
        a.W1, a.b1 = _p(w1), _p(bs[0])
        a.W_hidden = (C.c_void_p * 3)(*[w.data_ptr() for w in wh[:3]])
        a.b_hidden = (C.c_void_p * 3)(*[(b.data_ptr() if b is not None else None) for b in bs[1:4]])
        a.emb = _p(e)
        a.W5, a.b5 = _p(wh[3]), _p(bs[4])
        a.uv, a.jacobian = _p(uv), _p(jac)
        a.stash = (C.c_void_p * 4)(*[(s.data_ptr() if s is not None else None) for s in stash])
        a.stash_inv_len = _p(inv_len)
        a.debug_accumulators = _p(dbg)
        if N > 0:
            with torch.cuda.device(dev):
                L.check(lib.texgs_uvmlp_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "texgs_uvmlp_forward")
        ctx.offset, ctx.inv_scale = [float(v) for v in offset], [float(v) for v in inv_scale]
        ctx.has_bias = [b is not None for b in (b1, b2, b3, b4, b5)]
        if need_grad:
            ctx.save_for_backward(x, w1, *wh, uv, inv_len, *stash)
        ctx.mark_non_differentiable(jac)
        if debug:
            _UvMlp.last_debug = dbg
        return uv, jac

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_uv, _g_jac):
        """Per hidden layer ONE tcgen05 kernel (``texgs_uvmlp_backward_layer``: delta^T @ a into TMEM across the CTA's tiles with
        MN-major operands, delta @ W with K-major operands, ReLU mask and bias / embedding column sums in the epilogue; fp16
        operands on the stash the forward kernel wrote, fp32 accumulation); the K = 3 input and output layers and the
        normalisation are handled entirely by the ``head`` / ``tail`` kernels. No cuBLAS call is left on the path. The back-propagated signal stays in fp16 under one power-of-two scale chosen on
        the device from max|d loss / d out| (no host sync), as mixed-precision training does."""
        lib = L.load()
        saved = ctx.saved_tensors
        x, w1, w2h, w3h, w4h, w5h, uv, inv_len, a1, a2, a3, a4 = saved
        need = ctx.needs_input_grad
        dev = x.device
        N = x.shape[0]
        f32, f16 = torch.float32, torch.float16
        g = g_uv.to(f32).contiguous()
        z = lambda *shape: torch.zeros(*shape, device=dev, dtype=f32)
        gW5, gb5, gW1 = z(3, HIDDEN), z(3), z(HIDDEN, 3)
        cs = z(4, HIDDEN)                       # column sums of delta4..delta1 (scaled): bias / embedding gradients
        scratch, scale = z(1), z(1)
        d = torch.empty(N, HIDDEN, device=dev, dtype=f16)
        gx = torch.empty(N, 3, device=dev, dtype=f32) if need[0] else None
        if N == 0:
            zw = z(HIDDEN, HIDDEN)
            hb = ctx.has_bias
            return (gx, z(HIDDEN), gW1, z(HIDDEN) if hb[0] else None, zw, z(HIDDEN) if hb[1] else None, zw.clone(), z(HIDDEN) if hb[2] else None,
                    zw.clone(), z(HIDDEN) if hb[3] else None, gW5, gb5 if hb[4] else None, None, None, None)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            L.check(lib.texgs_uvmlp_backward_head(N, _p(g), _p(uv), _p(inv_len), _p(a4), _p(w5h), _p(scratch), _p(scale), _p(d), _p(gW5),
                                                  _p(gb5), _p(cs[0]), st), "texgs_uvmlp_backward_head")
            inv = 1.0 / scale
            # hidden layers 4, 3, 2: one tcgen05 kernel each — gW += delta^T a (K = points, MN-major operands), delta <- (delta W) *
            # (a > 0) (K-major operands, W^T resident in shared memory) and the column sums of the new delta
            d2 = torch.empty_like(d)
            gWs = []
            for w, a_prev, k in ((w4h, a3, 1), (w3h, a2, 2), (w2h, a1, 3)):
                gw = z(HIDDEN, HIDDEN)
                L.check(lib.texgs_uvmlp_backward_layer(N, _p(d), _p(a_prev), _p(w.T.contiguous()), _p(d2), _p(gw), _p(cs[k]), st),
                        "texgs_uvmlp_backward_layer")
                gWs.append(gw * inv)
                d, d2 = d2, d
            gW4, gW3, gW2 = gWs
            off3, isc3 = (C.c_float * 3)(*ctx.offset), (C.c_float * 3)(*ctx.inv_scale)
            L.check(lib.texgs_uvmlp_backward_tail(N, _p(d), _p(x), off3, isc3, _p(w1), _p(scale), _p(gx), _p(gW1), st), "texgs_uvmlp_backward_tail")
            cs = cs * inv
        gb4, gb3, gemb, gb1 = cs[0], cs[1], cs[2], cs[3]
        hb = ctx.has_bias
        return (gx, gemb, gW1, gb1 if hb[0] else None, gW2, gemb if hb[1] else None, gW3, gb3 if hb[2] else None,
                gW4, gb4 if hb[3] else None, gW5, gb5 if hb[4] else None, None, None, None)


class FusedUVNet(nn.Module):
    """Drop-in for the reference's ``UVNet`` (``models/modules/uv_net.py:8-36``) with the 128-wide configuration of
    ``configs/texture_gaussian3d.yaml:18-27``. ``forward(xyz, emb)`` returns uv like the reference;
    ``uv_and_jacobian(xyz, emb)`` returns ``(uv, grad_uvs)`` — what ``get_uvs`` and ``get_grad_uvs`` return."""

    def __init__(self, bias: bool = True, xyz_offset: Optional[Sequence[float]] = None, xyz_scale: Optional[Sequence[float]] = None):
        super().__init__()
        self.pre_mlp = nn.Sequential(nn.Linear(3, HIDDEN, bias=bias), nn.ReLU(), nn.Linear(HIDDEN, HIDDEN, bias=bias))
        self.mlp = nn.Sequential(nn.Linear(HIDDEN, HIDDEN, bias=bias), nn.ReLU(), nn.Linear(HIDDEN, HIDDEN, bias=bias), nn.ReLU(),
                                 nn.Linear(HIDDEN, 3, bias=bias))
        if (xyz_offset is None) != (xyz_scale is None):
            raise ValueError("xyz_offset and xyz_scale come together (models/modules/uv_net.py:22-25)")
        self.xyz_offset = None if xyz_offset is None else [float(v) for v in xyz_offset]
        self.xyz_scale = None if xyz_scale is None else [float(v) for v in xyz_scale]

    def _call(self, xyz, emb, debug=None):
        off = self.xyz_offset or [0.0, 0.0, 0.0]
        isc = [1.0 / v for v in self.xyz_scale] if self.xyz_scale else [1.0, 1.0, 1.0]
        l = (self.pre_mlp[0], self.pre_mlp[2], self.mlp[0], self.mlp[2], self.mlp[4])
        return _UvMlp.apply(xyz, emb, l[0].weight, l[0].bias, l[1].weight, l[1].bias, l[2].weight, l[2].bias, l[3].weight, l[3].bias,
                            l[4].weight, l[4].bias, off, isc, debug)

    def forward(self, xyz: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
        return self._call(xyz, emb)[0]

    def uv_and_jacobian(self, xyz: torch.Tensor, emb: torch.Tensor):
        """``(uv (N,3), grad_uvs (N,9))``; ``grad_uvs[n, 3i+j] = d uv_i / d xyz_j``, no gradient (reference ``:227``)."""
        return self._call(xyz, emb)
