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

    def __init__(self, params: Dict[str, torch.Tensor], pad_texture: bool = True, symmetric_group=None):
        """One flat buffer for all streams: every accumulation the backward kernels do into it is atomic (vector
        ``red`` for texels, ``red`` / TMA bulk reductions for the per-Gaussian gradients), so views rendered
        concurrently on several CUDA streams (``render_views_accumulate(..., streams=n)``) share it.

        ``symmetric_group`` (a process group, CUDA + NCCL only): the buffer is allocated as symmetric memory and
        exchanged with the group's ranks, so that their kernels can read it over NVLink (``DistTextureAdam``: the
        texture gradient is then reduced by the owner ranks instead of being all-reduced)."""
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
        self.symm = None
        if symmetric_group is not None:
            import torch.distributed._symmetric_memory as symm_mem
            self.flat = symm_mem.empty(off, dtype=torch.float32, device=first.device)
            self.flat.zero_()
            self.symm = symm_mem.rendezvous(self.flat, symmetric_group.group_name)
        else:
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
        flat = self.flat
        if self.padded[k]:
            return flat[o:o + n].view(*v.shape[:-1], 4), True
        return flat[o:o + n].view_as(v), False

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
        """Rasterizer calls made inside this block accumulate their gradients into the bucket from within the backward
        kernels (for inputs that ARE bucket leaves; others go through autograd)."""
        prev = getattr(_fused, "bucket", None)
        _fused.bucket = self
        try:
            yield self
        finally:
            _fused.bucket = prev

    def all_reduce(self, group=None, async_op: bool = False, exclude: Sequence[str] = ()):
        """Sum over ranks (one NCCL all-reduce per step over the flat bucket). ``exclude``: parameter names
        whose slices are left alone (the texture when ``DistTextureAdam`` reduces it itself); the rest goes out as the
        contiguous ranges around them. Returns the work handle(s) with ``async_op``."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        if not exclude:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        works = [dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
                 for a, b in self.ranges_without(exclude)]
        return works if async_op else None

    def ranges_without(self, exclude: Sequence[str]):
        """Contiguous [a, b) element ranges of the flat buffer that do not belong to the named parameters."""
        cuts = sorted((self.offsets[k][0], self.offsets[k][0] + (self.offsets[k][1] + 63) // 64 * 64) for k in exclude if k in self.offsets)
        out, pos = [], 0
        for a, b in cuts:
            if a > pos:
                out.append((pos, a))
            pos = max(pos, b)
        if pos < self.flat.numel():
            out.append((pos, self.flat.numel()))
        return out

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4


class DistTextureAdam:
    """The texture's optimizer step across ``world`` GPUs without an all-reduce of its gradient.

    Every rank accumulates the texture gradient of its views into a symmetric ``GradBucket``. ``step()`` then runs ONE
    kernel per rank (``texgs_texture_adam_dp_step``): the rank owns 1/world of the texels, pulls the ``world`` partial
    gradients of those texels over NVLink — ``multimem.ld_reduce`` through the NVSwitch multicast mapping when the
    fabric offers one, plain peer loads otherwise —, applies ``torch.optim.Adam``'s update (the reference's optimizer,
    ``models/texture_gaussian3d.py:139-143``) to its shard of the moments and pushes the updated texels into every
    rank's copy of the texture (``multimem.st`` / peer stores). Device-side barriers over the symmetric-memory signal
    pads bracket the kernel. The texture's storage is moved into symmetric memory at construction (same tensor object);
    the Adam moments exist for the owned shard only (``state_shard()``).

    Same hyper-parameter semantics as ``TextureAdam`` / ``torch.optim.Adam``; ``param_groups[0]['lr']`` may be changed
    between steps (schedulers)."""

    def __init__(self, texture: torch.Tensor, bucket: "GradBucket", lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 group=None, name: str = "texture", use_multicast: Optional[bool] = None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        if bucket.symm is None:
            raise ValueError("DistTextureAdam needs a GradBucket(..., symmetric_group=group)")
        if not (texture.is_cuda and texture.dtype == torch.float32 and texture.is_contiguous() and texture.numel() % 3 == 0):
            raise L.TexgsError("DistTextureAdam: the texture must be a contiguous fp32 CUDA tensor with numel % 3 == 0")
        if bucket.storage_for(texture) is None or not bucket.padded[name]:
            raise ValueError("the texture must be a (padded) leaf of the bucket")
        group = group or dist.group.WORLD
        self.group, self.bucket, self.texture, self.name = group, bucket, texture, name
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.param_groups = [dict(lr=lr, betas=tuple(betas), eps=eps)]
        self.n_texels = texture.numel() // 3
        # the parameter moves into symmetric memory: peers write their shards of the update into it
        buf = symm_mem.empty(texture.numel(), dtype=torch.float32, device=texture.device)
        buf.copy_(texture.detach().reshape(-1))
        texture.data = buf.view(texture.shape)
        self.param_symm = symm_mem.rendezvous(buf, group.group_name)
        lo, hi = C.c_uint64(), C.c_uint64()
        L.check(L.load().texgs_dp_shard(self.n_texels, self.world, self.rank, C.byref(lo), C.byref(hi)), "texgs_dp_shard")
        self.tile_lo, self.tile_hi = lo.value, hi.value
        own = max(1, (self.tile_hi - self.tile_lo) * 1024 * 3)
        self.exp_avg = torch.zeros(own, dtype=torch.float32, device=texture.device)
        self.exp_avg_sq = torch.zeros(own, dtype=torch.float32, device=texture.device)
        self.step_count = 0
        g_mc, p_mc = int(getattr(bucket.symm, "multicast_ptr", 0) or 0), int(getattr(self.param_symm, "multicast_ptr", 0) or 0)
        self.multicast = bool(g_mc and p_mc) if use_multicast is None else bool(use_multicast and g_mc and p_mc)
        self._mc = (g_mc, p_mc)

    def state_shard(self):
        """(first owned texel, exp_avg, exp_avg_sq) — the moments of texels [tile_lo*1024, min(tile_hi*1024, n))."""
        return self.tile_lo * 1024, self.exp_avg, self.exp_avg_sq

    @torch.no_grad()
    def step(self):
        import ctypes as C
        from . import _lib as L
        lib = L.load()
        dev = self.texture.device
        self.step_count += 1
        pg = self.param_groups[0]
        a = L.TexgsDpAdamArgs()
        a.world, a.rank = self.world, self.rank
        off = 4 * self.bucket.offsets[self.name][0]
        for r in range(self.world):
            a.grad_ptrs[r] = int(self.bucket.symm.buffer_ptrs[r]) + off
            a.param_ptrs[r] = int(self.param_symm.buffer_ptrs[r])
        if self.multicast:
            a.grad_mc, a.param_mc = self._mc[0] + off, self._mc[1]
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.n_texels, a.tile_lo, a.tile_hi = self.n_texels, self.tile_lo, self.tile_hi
        a.lr, (a.beta1, a.beta2), a.eps, a.step = float(pg["lr"]), pg["betas"], float(pg["eps"]), self.step_count
        with torch.cuda.device(dev):
            self.bucket.symm.barrier(channel=0)            # every rank's backward has landed in its gradient buffer
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            L.check(lib.texgs_texture_adam_dp_step(C.byref(a), st), "texgs_texture_adam_dp_step")
            self.bucket.symm.barrier(channel=1)            # every owner is done reading gradients / writing texels
        torch.autograd.graph.increment_version(self.texture)   # written through raw pointers (the packed copy is rebuilt)


_stream_pools: dict = {}


def _stream_pool(device: torch.device, n: int):
    key = (device.index, n)
    if key not in _stream_pools:
        _stream_pools[key] = [torch.cuda.Stream(device) for _ in range(n)]
    return _stream_pools[key]


def render_views_accumulate(render_fn, gaussians, cameras: Sequence, cotangents, view_ids: Iterable[int], bg,
                            timer=None, bucket: Optional["GradBucket"] = None, streams: int = 1, backward: bool = True,
                            loss_fn=None, before_view=None, after_view=None, fork_event=None, render_event=None):
    """Forward (+ backward) of ``render_fn`` (``uv_tex_render``) for the given views; gradients accumulate into the
    leaves' ``.grad`` (i.e. the bucket) — from inside the backward kernels when ``bucket`` is given
    (``bucket.fused()``), through autograd otherwise.

    What drives the backward: fixed dense output cotangents (``cotangents``: a 4-tuple for render / depth / norm /
    alpha, or a callable view -> 4-tuple), or ``loss_fn(pkg, view) -> scalar`` (the training-shaped step: losses on
    the rasterizer outputs, ``models/texture_gaussian3d.py:333-368``). ``backward=False`` renders only
    (``retexture.py:18-37``). ``before_view(view, slot)`` / ``after_view(view, slot)`` run on the stream that renders
    the view (stream waits / event records of a host-fed input pipeline); ``slot`` counts the views of this call.

    ``streams`` > 1 (needs a bucket whose leaves are ALL the differentiable inputs when ``backward``): view i runs on
    CUDA stream i mod ``streams``, so the latency-bound small kernels of one view (tile scan, scatter, sort) and the
    tails of its render kernels overlap the render kernels of the next one (+9 % views/s at the headline size,
    profiles/r2_variants.md). All streams accumulate into the one bucket (atomic adds). The packed texel copy is built
    once before the fork; the calling stream joins all streams before returning.

    ``fork_event`` / ``render_event`` (``streams`` > 1) overlap the tail of the PREVIOUS batch with the head of this
    one: the view streams start after ``fork_event`` (e.g. recorded when the previous batch's views had joined) instead
    of after everything the calling stream has queued since (gradient reduction, optimizer step, repack, bucket clear),
    and only the render kernel of each view — the first to touch the texture; the backward, the first to touch the
    bucket, comes after it — waits for ``render_event``, recorded behind that work."""
    view_ids = list(view_ids)
    from .rasterizer import wait_before_render

    def one_view(slot, v, fused_ctx):
        cam = cameras[v % len(cameras)]
        ctx = timer.view(backward=backward) if timer is not None else _null()
        with ctx, fused_ctx, wait_before_render(render_event):
            if before_view is not None:
                before_view(v, slot)
            if not backward:
                with torch.no_grad():
                    render_fn(cam, gaussians, None, bg)
            else:
                pkg = render_fn(cam, gaussians, None, bg)
                if loss_fn is not None:
                    loss_fn(pkg, v).backward()
                else:
                    cot = cotangents(v) if callable(cotangents) else cotangents
                    torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
            if after_view is not None:
                after_view(v, slot)

    if streams <= 1 or len(view_ids) <= 1:
        for i, v in enumerate(view_ids):
            one_view(i, v, bucket.fused() if (bucket is not None and backward) else _null())
        return len(view_ids)
    if backward:
        if bucket is None:
            raise ValueError("streams > 1 needs a GradBucket: autograd would add the views' gradients into one .grad from several streams at once")
        # autograd would add the gradients of non-bucket leaves into ONE .grad tensor from several streams at once
        for t in _differentiable_inputs(gaussians):
            if t.requires_grad and (not t.is_leaf or bucket.storage_for(t) is None):
                raise ValueError("streams > 1: every differentiable input of the render must be a leaf of the bucket "
                                 "(activations computed outside it would be accumulated by autograd from several streams at once)")
    dev = bucket.flat.device if bucket is not None else torch.device("cuda", torch.cuda.current_device())
    from .rasterizer import ensure_packed_texture, packed_is_current
    tex = getattr(gaussians, "get_texture", None)
    early = fork_event is not None and render_event is not None
    if tex is not None:
        if early and not packed_is_current(tex):
            early = False                                # the repack below is newer than render_event: plain fork
        ensure_packed_texture(tex)                       # on the calling stream, before the fork
    main = torch.cuda.current_stream(dev)
    pool = _stream_pool(dev, streams)
    for s in pool:
        if early:
            s.wait_event(fork_event)
        else:
            s.wait_stream(main)
    for i, v in enumerate(view_ids):
        r = i % streams
        with torch.cuda.stream(pool[r]):
            one_view(i, v, bucket.fused() if (bucket is not None and backward) else _null())
    for s in pool:
        main.wait_stream(s)
    return len(view_ids)


def _differentiable_inputs(gaussians):
    out = []
    for name in ("get_xyz", "get_opacity", "get_scaling", "get_rotation", "get_shs", "get_texture", "get_uvs"):
        t = getattr(gaussians, name, None)
        if isinstance(t, torch.Tensor):
            out.append(t)
    return out


def bind_to_gpu_numa_node(device: torch.device) -> Optional[int]:
    """Restrict the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that pinned host buffers
    allocated afterwards are first-touched on that node and the host-to-device copies of a rank do not cross the
    socket interconnect (round 1: eight ranks feeding from node 0 was what bent the end-to-end scaling curve).
    Returns the node, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
