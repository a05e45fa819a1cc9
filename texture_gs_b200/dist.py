"""Camera-batch data parallelism: the only multi-GPU axis of this path.

Gaussians + texture are replicated on every rank, the view batch is sharded, every rank accumulates
the gradients of its local views into ONE flat fp32 bucket (the leaves' ``.grad`` tensors are views
into it, so the rasterizer backward / autograd accumulate straight into the communication buffer —
no gather/copy before the collective), and a single all-reduce(sum) per step makes the bucket
identical on all ranks. One process per GPU; backend nccl on GPUs (NVLink 5 / NVSwitch), gloo on
CPU for the host-logic tests.
"""
from __future__ import annotations

import threading
from contextlib import contextmanager
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def init_process_group_quiet(backend: str, device: Optional[torch.device] = None, **kw) -> None:
    """``dist.init_process_group`` + one barrier with file descriptor 1 pointed at stderr meanwhile.

    NCCL announces its version on stdout while the communicator is created (at init when ``device_id`` is given,
    else at the first collective). A launcher whose stdout carries a machine-readable result — ``bench.py`` prints
    exactly one JSON line — must not have that line in front of it. Everything printed after this call goes to the
    real stdout again."""
    import os
    import sys
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        if device is not None and device.type == "cuda":
            dist.init_process_group(backend, device_id=device, **kw)
        else:
            dist.init_process_group(backend, **kw)
        dist.barrier()
        if device is not None and device.type == "cuda":
            torch.cuda.synchronize(device)
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)


def shard_views(num_views: int, world_size: int, rank: int) -> List[int]:
    """Indices of the views rank ``rank`` renders: contiguous blocks, sizes differ by at most one."""
    base, rem = divmod(num_views, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


_fused = threading.local()


def current_fused_bucket():
    """The GradBucket of the calling thread's active ``bucket.fused()`` block, or None."""
    return getattr(_fused, "bucket", None)


class GradBucket:
    """Flat gradient buffer whose slices are installed as ``.grad`` of the given leaf tensors.

    A ``(6,R,R,3)`` texture is stored PADDED to 4 floats per texel inside the bucket; its ``.grad``
    is the strided ``[..., :3]`` view. That is the layout the backward kernel's 128-bit vector
    atomics (``red.global.add.v4.f32``) write, so inside ``with bucket.fused():`` the rasterizer
    backward accumulates texture AND per-Gaussian gradients directly into this buffer — the buffer
    the all-reduce runs on — and autograd's separate ``AccumulateGrad`` pass (a zero-fill, a dense
    gradient tensor and an add kernel per input per view) disappears."""

    def __init__(self, params: Dict[str, torch.Tensor], pad_texture: bool = True):
        self.params = {k: v for k, v in params.items() if v is not None and v.requires_grad}
        if not self.params:
            raise ValueError("no tensor requires grad")
        first = next(iter(self.params.values()))
        self.offsets, self.padded = {}, {}
        off = 0
        for k, v in self.params.items():
            pad = bool(pad_texture and v.dim() == 4 and v.shape[0] == 6 and v.shape[-1] == 3 and v.shape[1] == v.shape[2])
            n = v.numel() // 3 * 4 if pad else v.numel()
            self.offsets[k] = (off, n)
            self.padded[k] = pad
            off += (n + 63) // 64 * 64          # keep every slice 256-byte aligned
        self.flat = torch.zeros(off, dtype=torch.float32, device=first.device)
        self._by_id = {id(v): k for k, v in self.params.items()}
        self.install()

    def _view(self, k):
        o, n = self.offsets[k]
        v = self.params[k]
        if self.padded[k]:
            return self.flat[o:o + n].view(*v.shape[:-1], 4)[..., :3]
        return self.flat[o:o + n].view_as(v)

    def storage_for(self, tensor: torch.Tensor):
        """(buffer, padded) the rasterizer may accumulate into for this exact leaf tensor, else None."""
        k = self._by_id.get(id(tensor))
        if k is None or self.params[k] is not tensor:
            return None
        o, n = self.offsets[k]
        v = self.params[k]
        if self.padded[k]:
            return self.flat[o:o + n].view(*v.shape[:-1], 4), True
        return self.flat[o:o + n].view_as(v), False

    def install(self):
        for k, v in self.params.items():
            v.grad = self._view(k)

    def zero(self):
        self.flat.zero_()
        for k, v in self.params.items():       # autograd keeps a defined .grad in place; re-check cheaply
            o, _ = self.offsets[k]
            if v.grad is None or v.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
                v.grad = self._view(k)

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: self._view(k) for k in self.params}

    @contextmanager
    def fused(self):
        """Rasterizer calls made inside this block accumulate their gradients into the bucket from
        within the backward kernels (for inputs that ARE bucket leaves; others go through autograd)."""
        prev = getattr(_fused, "bucket", None)
        _fused.bucket = self
        try:
            yield self
        finally:
            _fused.bucket = prev

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum over ranks (one NCCL all-reduce per step over the flat bucket)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4


def render_views_accumulate(render_fn, gaussians, cameras: Sequence, cotangents, view_ids: Iterable[int], bg,
                            timer=None, bucket: Optional["GradBucket"] = None):
    """Forward + backward of ``render_fn`` (``uv_tex_render``) for the given views with fixed dense
    output cotangents; gradients accumulate into the leaves' ``.grad`` (i.e. the bucket) — from inside
    the backward kernels when ``bucket`` is given (``bucket.fused()``), through autograd otherwise."""
    n = 0
    for v in view_ids:
        cam = cameras[v % len(cameras)]
        cot = cotangents(v) if callable(cotangents) else cotangents
        ctx = timer.view() if timer is not None else _null()
        with ctx, (bucket.fused() if bucket is not None else _null()):
            pkg = render_fn(cam, gaussians, None, bg)
            torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
        n += 1
    return n


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
