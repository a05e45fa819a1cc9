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

    def __init__(self, params: Dict[str, torch.Tensor], pad_texture: bool = True, replicas: int = 1):
        """``replicas`` > 1 adds private copies of the flat buffer, one per CUDA stream of
        ``render_views_accumulate(..., streams=replicas)``: views rendered concurrently on different streams must
        not share an accumulation buffer (the per-Gaussian backward accumulates with plain read-modify-writes);
        ``reduce_replicas()`` — called by ``all_reduce()`` — folds them into replica 0, the ``.grad`` storage."""
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
        self.flats = [self.flat] + [torch.zeros_like(self.flat) for _ in range(max(1, int(replicas)) - 1)]
        self._by_id = {id(v): k for k, v in self.params.items()}
        self.install()

    def _view(self, k):
        o, n = self.offsets[k]
        v = self.params[k]
        if self.padded[k]:
            return self.flat[o:o + n].view(*v.shape[:-1], 4)[..., :3]
        return self.flat[o:o + n].view_as(v)

    def storage_for(self, tensor: torch.Tensor):
        """(buffer, padded) the rasterizer may accumulate into for this exact leaf tensor, else None. Inside
        ``fused(replica=r)`` the buffer is replica ``r``'s."""
        k = self._by_id.get(id(tensor))
        if k is None or self.params[k] is not tensor:
            return None
        o, n = self.offsets[k]
        v = self.params[k]
        r = getattr(_fused, "replica", 0) if getattr(_fused, "bucket", None) is self else 0
        flat = self.flats[r]
        if self.padded[k]:
            return flat[o:o + n].view(*v.shape[:-1], 4), True
        return flat[o:o + n].view_as(v), False

    def install(self):
        for k, v in self.params.items():
            v.grad = self._view(k)

    def zero(self):
        for f in self.flats:
            f.zero_()
        for k, v in self.params.items():       # autograd keeps a defined .grad in place; re-check cheaply
            o, _ = self.offsets[k]
            if v.grad is None or v.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
                v.grad = self._view(k)

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: self._view(k) for k in self.params}

    @contextmanager
    def fused(self, replica: int = 0):
        """Rasterizer calls made inside this block accumulate their gradients into the bucket (its replica
        ``replica``) from within the backward kernels (for inputs that ARE bucket leaves; others go through autograd)."""
        if not 0 <= replica < len(self.flats):
            raise ValueError(f"replica {replica} of a bucket with {len(self.flats)}")
        prev = getattr(_fused, "bucket", None), getattr(_fused, "replica", 0)
        _fused.bucket, _fused.replica = self, replica
        try:
            yield self
        finally:
            _fused.bucket, _fused.replica = prev

    def reduce_replicas(self):
        """Fold the per-stream replicas into replica 0 (the ``.grad`` storage) and clear them; call on a stream that
        has waited for every stream that wrote a replica."""
        for f in self.flats[1:]:
            self.flat.add_(f)
            f.zero_()

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum over ranks (one NCCL all-reduce per step over the flat bucket)."""
        self.reduce_replicas()
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4


_stream_pools: dict = {}


def _stream_pool(device: torch.device, n: int):
    key = (device.index, n)
    if key not in _stream_pools:
        _stream_pools[key] = [torch.cuda.Stream(device) for _ in range(n)]
    return _stream_pools[key]


def render_views_accumulate(render_fn, gaussians, cameras: Sequence, cotangents, view_ids: Iterable[int], bg,
                            timer=None, bucket: Optional["GradBucket"] = None, streams: int = 1):
    """Forward + backward of ``render_fn`` (``uv_tex_render``) for the given views with fixed dense
    output cotangents; gradients accumulate into the leaves' ``.grad`` (i.e. the bucket) — from inside
    the backward kernels when ``bucket`` is given (``bucket.fused()``), through autograd otherwise.

    ``streams`` > 1 (experimental; needs a bucket with as many replicas whose leaves are ALL the differentiable
    inputs): view i runs on CUDA stream i mod ``streams`` and accumulates into that stream's replica, so the
    latency-bound small kernels of one view (tile scan, scatter, sort) and the tails of its render kernels overlap
    the render kernels of the next one. The packed texel copy is built once before the fork; the calling stream
    joins all streams before returning, the replicas are folded by ``bucket.all_reduce()`` / ``reduce_replicas()``."""
    view_ids = list(view_ids)
    if streams <= 1 or len(view_ids) <= 1:
        for v in view_ids:
            cam = cameras[v % len(cameras)]
            cot = cotangents(v) if callable(cotangents) else cotangents
            ctx = timer.view() if timer is not None else _null()
            with ctx, (bucket.fused() if bucket is not None else _null()):
                pkg = render_fn(cam, gaussians, None, bg)
                torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
        return len(view_ids)
    if bucket is None or len(bucket.flats) < streams:
        raise ValueError("streams > 1 needs a GradBucket(..., replicas=streams): concurrent views must not share gradient buffers")
    dev = bucket.flat.device
    from .rasterizer import ensure_packed_texture
    tex = getattr(gaussians, "get_texture", None)
    if tex is not None:
        ensure_packed_texture(tex)                       # on the calling stream, before the fork
    main = torch.cuda.current_stream(dev)
    pool = _stream_pool(dev, streams)
    for s in pool:
        s.wait_stream(main)
    for i, v in enumerate(view_ids):
        r = i % streams
        cam = cameras[v % len(cameras)]
        cot = cotangents(v) if callable(cotangents) else cotangents
        ctx = timer.view() if timer is not None else _null()
        with torch.cuda.stream(pool[r]), ctx, bucket.fused(replica=r):
            pkg = render_fn(cam, gaussians, None, bg)
            torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
    for s in pool:
        main.wait_stream(s)
    return len(view_ids)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
