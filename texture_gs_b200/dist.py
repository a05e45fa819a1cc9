"""Camera-batch data parallelism: the only multi-GPU axis of this path.

Gaussians + texture are replicated on every rank, the view batch is sharded, every rank accumulates
the gradients of its local views into ONE flat fp32 bucket (the leaves' ``.grad`` tensors are views
into it, so the rasterizer backward / autograd accumulate straight into the communication buffer —
no gather/copy before the collective), and a single all-reduce(sum) per step makes the bucket
identical on all ranks. One process per GPU; backend nccl on GPUs (NVLink 5 / NVSwitch), gloo on
CPU for the host-logic tests.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, world_size: int, rank: int) -> List[int]:
    """Indices of the views rank ``rank`` renders: contiguous blocks, sizes differ by at most one."""
    base, rem = divmod(num_views, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


class GradBucket:
    """Flat gradient buffer whose slices are installed as ``.grad`` of the given leaf tensors."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.params = {k: v for k, v in params.items() if v is not None and v.requires_grad}
        if not self.params:
            raise ValueError("no tensor requires grad")
        first = next(iter(self.params.values()))
        self.offsets = {}
        off = 0
        for k, v in self.params.items():
            self.offsets[k] = (off, v.numel())
            off += (v.numel() + 63) // 64 * 64          # keep every slice 256-byte aligned
        self.flat = torch.zeros(off, dtype=torch.float32, device=first.device)
        self.install()

    def install(self):
        for k, v in self.params.items():
            o, n = self.offsets[k]
            v.grad = self.flat[o:o + n].view_as(v)

    def zero(self):
        self.flat.zero_()
        # autograd may have replaced .grad (it does not when .grad is already defined), re-check cheaply
        for k, v in self.params.items():
            o, n = self.offsets[k]
            if v.grad is None or v.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
                v.grad = self.flat[o:o + n].view_as(v)

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: self.flat[o:o + n].view_as(self.params[k]) for k, (o, n) in self.offsets.items()}

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum over ranks (one NCCL all-reduce per step over the flat bucket)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4


def render_views_accumulate(render_fn, gaussians, cameras: Sequence, cotangents, view_ids: Iterable[int], bg,
                            timer=None):
    """Forward + backward of ``render_fn`` (``uv_tex_render``) for the given views with fixed dense
    output cotangents; gradients accumulate into the leaves' ``.grad`` (i.e. the bucket)."""
    n = 0
    for v in view_ids:
        cam = cameras[v % len(cameras)]
        cot = cotangents(v) if callable(cotangents) else cotangents
        ctx = timer.view() if timer is not None else _null()
        with ctx:
            pkg = render_fn(cam, gaussians, None, bg)
            torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
        n += 1
    return n


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
