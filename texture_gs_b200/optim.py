"""SURVEY §8f N4 — the texture's optimizer step, fused with the buffers around it on this path.

``TextureAdam`` is a drop-in for the ``torch.optim.Adam(l, lr=0.0, eps=1e-15)`` the reference builds over the
``(6,R,R,3)`` texture (``models/texture_gaussian3d.py:139-143``) and steps every iteration (``:439-440``): same
constructor arguments, same ``param_groups`` keys (so ``param_group['lr'] = ...`` schedules keep working), same
``state`` layout (``step`` / ``exp_avg`` / ``exp_avg_sq``), so ``state_dict()`` / ``load_state_dict()`` interchange
with the reference's checkpoints (``:150-155, :191-193``).

One kernel per parameter does the dense Adam update and, in the same pass,
  * reads the gradient directly from the padded ``(6,R,R,4)`` storage a ``GradBucket`` gives the texture (the buffer
    the rasterizer backward's vector atomics accumulate into) — or from a plain contiguous ``.grad``;
  * optionally clears that gradient for the next step (``zero_grad_in_step=True``: replaces the 400 MB fill);
  * writes the packed RGBA copy of the UPDATED texture and hands it to the rasterizer's cache, so the next forward
    does not repack.
No CPU fallback: parameters must be contiguous fp32 CUDA tensors with ``numel % 3 == 0``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _is_cube_texture(p: torch.Tensor) -> bool:
    return p.dim() == 4 and p.shape[0] == 6 and p.shape[1] == p.shape[2] and p.shape[3] == 3


def _padded_storage(grad: torch.Tensor):
    """If ``grad`` is the ``[..., :3]`` view of a contiguous ``(..., 4)`` buffer, return that buffer's data pointer."""
    if grad.dim() < 1 or grad.shape[-1] != 3 or grad.stride(-1) != 1:
        return None
    expect = 4
    for size, stride in zip(reversed(grad.shape[:-1]), reversed(grad.stride()[:-1])):
        if size != 1 and stride != expect:
            return None
        expect *= size
    return grad.data_ptr()


class TextureAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False, *, maximize: bool = False, zero_grad_in_step: bool = False,
                 emit_packed_texture: bool = True):
        if weight_decay != 0.0 or amsgrad or maximize:
            raise NotImplementedError("TextureAdam implements the plain Adam the reference uses (no weight decay / amsgrad / maximize)")
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0.0, amsgrad=False, maximize=False)
        super().__init__(params, defaults)
        self.zero_grad_in_step = bool(zero_grad_in_step)
        self.emit_packed_texture = bool(emit_packed_texture)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.numel() % 3 == 0):
                    raise L.TexgsError("TextureAdam: parameters must be contiguous fp32 CUDA tensors with numel % 3 == 0 "
                                       f"(no fallback); got {tuple(p.shape)} {p.dtype} on {p.device}")
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32 or g.device != p.device:
                    raise L.TexgsError("TextureAdam: dense fp32 gradients on the parameter's device only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                step = int(st["step"].item()) if st["step"].device.type == "cpu" else int(st["step"])
                g4 = _padded_storage(g) if not g.is_contiguous() else None
                g3 = None
                if g4 is None:
                    g = g if g.is_contiguous() else g.contiguous()
                    g3 = g.data_ptr()
                rgba = None
                if self.emit_packed_texture and _is_cube_texture(p):
                    from . import rasterizer as R
                    rgba = R.packed_texture_buffer(p)
                with torch.cuda.device(p.device):
                    stream = torch.cuda.current_stream(p.device).cuda_stream
                    L.check(lib.texgs_texture_adam_step(
                        C.c_void_p(p.data_ptr()), C.c_void_p(st["exp_avg"].data_ptr()), C.c_void_p(st["exp_avg_sq"].data_ptr()),
                        C.c_void_p(g3) if g3 is not None else None, C.c_void_p(g4) if g4 is not None else None,
                        C.c_void_p(rgba.data_ptr()) if rgba is not None else None, p.numel() // 3,
                        float(group["lr"]), float(b1), float(b2), float(group["eps"]), step,
                        1 if self.zero_grad_in_step else 0, C.c_void_p(stream)), "texgs_texture_adam_step")
                torch.autograd.graph.increment_version(p)        # written through a raw pointer
                if rgba is not None:
                    R.adopt_packed_texture(p, rgba)
        return loss
