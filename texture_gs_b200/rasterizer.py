"""Host side of the drop-in boundary: the Python surface of ``diff_gauss_uv_tex`` / ``diff_gauss``.

Mirrors what the reference imports at ``render/uv_tex_render.py:4`` and ``render/render.py:4``:

* ``GaussianRasterizationSettings`` — the 12 fields built at ``render/uv_tex_render.py:25-38``;
* ``GaussianRasterizer(raster_settings=...)`` — an ``nn.Module`` called with the kwargs of
  ``render/uv_tex_render.py:56-66`` (textured) or ``render/render.py:75-84`` (plain 3DGS) and
  returning ``(image, depth, norm, alpha, radii, extra)``.

Underneath: one ``torch.autograd.Function`` whose forward/backward each make ONE call into the
C-ABI library (``include/texgs.h``) with raw device pointers on the current CUDA stream.
PyTorch owns every buffer. There is no CPU / eager fallback: without libtexgs.so or without a CUDA
device every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref
from contextlib import contextmanager
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _lib as L


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    debug: bool = False


# ---------------------------------------------------------------------------------------------
# small host-side caches (no device state)
# ---------------------------------------------------------------------------------------------

_host_cache: dict = {}
_host_cache_lock = threading.Lock()


def _host_floats(t: torch.Tensor, n: int):
    """Host copy of a small settings tensor (view/proj matrices, campos, bg).

    The reference keeps these on the GPU (``utils/cameras.py:62-65``); reading them costs one
    device->host sync the first time a given tensor object is seen, afterwards the copy is served
    from a cache keyed on object identity + storage address + in-place version counter. Writes that bypass the
    version counter (``t.data.copy_(...)``, raw-pointer updates from another library) are NOT seen: call
    ``invalidate_settings_cache()`` after such a write (cameras are immutable in the reference, ``utils/cameras.py``)."""
    if not isinstance(t, torch.Tensor):
        vals = [float(v) for v in t]
        assert len(vals) == n
        return vals
    if t.device.type == "cpu":
        vals = t.detach().reshape(-1).to(torch.float32).tolist()
        assert len(vals) == n, f"expected {n} values, got {len(vals)}"
        return vals
    key = id(t)
    with _host_cache_lock:
        hit = _host_cache.get(key)
        if hit is not None and hit[0]() is t and hit[1] == (t._version, t.data_ptr()):
            return hit[2]
    vals = t.detach().reshape(-1).to(torch.float32).cpu().tolist()
    assert len(vals) == n, f"expected {n} values, got {len(vals)}"
    with _host_cache_lock:
        if len(_host_cache) > 4096:
            _host_cache.clear()
        try:
            ref = weakref.ref(t, lambda _r, k=key: _host_cache.pop(k, None))
        except TypeError:
            return vals
        _host_cache[key] = (ref, (t._version, t.data_ptr()), vals)
    return vals


_capacity_hint: dict = {}          # (device index, P, H, W) -> pair capacity that last sufficed
_pinned: dict = {}                 # device index -> free list of (pinned int32[8], cuda event)
_pinned_lock = threading.Lock()


def _acquire_counters(dev_index: int):
    """A private (pinned counter buffer, event) pair for ONE forward call: the lock only guards the free list, never a
    wait, so forwards on different threads / streams of a device do not serialise on each other."""
    with _pinned_lock:
        free = _pinned.setdefault(dev_index, [])
        if free:
            return free.pop()
    buf = torch.zeros(8, dtype=torch.int32).pin_memory()
    ev = torch.cuda.Event(enable_timing=False, blocking=False)
    with torch.cuda.device(dev_index):
        ev.record()          # torch creates the cudaEvent_t lazily; force it so .cuda_event is valid
        ev.synchronize()
    return buf, ev


def _release_counters(dev_index: int, ent) -> None:
    with _pinned_lock:
        _pinned.setdefault(dev_index, []).append(ent)


# Packed (6,R,R,4) copy of a texture, cached per texture tensor object + in-place version: the
# optimizer step bumps ``_version`` so the copy is rebuilt exactly once per texture update (the
# reference renders 1-2 times per update, models/texture_gaussian3d.py:318,378; a view batch more).
USE_PACKED_TEXTURE = True
_packed_cache: dict = {}


def invalidate_packed_cache():
    """Forget every packed texel copy. Needed after a write to a texture that does not bump its version counter
    (``texture.data.clamp_()``, a raw-pointer update); in-place ops and optimizer steps are seen without it."""
    _packed_cache.clear()


def invalidate_settings_cache():
    """Forget the host copies of camera matrices / campos / bg (see ``_host_floats``)."""
    with _host_cache_lock:
        _host_cache.clear()


def _packed_texture(lib, texture_arg: torch.Tensor, tex: torch.Tensor, stream: int) -> torch.Tensor:
    key = id(texture_arg)
    hit = _packed_cache.get(key)
    if hit is not None and hit[0]() is texture_arg and hit[1] == texture_arg._version and hit[2] == tex.data_ptr():
        return hit[3]
    tex4 = torch.empty(tex.shape[0], tex.shape[1], tex.shape[2], 4, device=tex.device, dtype=torch.float32)
    L.check(lib.texgs_pack_texture(_ptr(tex), tex.shape[1], _ptr(tex4), C.c_void_p(stream)), "texgs_pack_texture")
    try:
        ref = weakref.ref(texture_arg, lambda _r, k=key: _packed_cache.pop(k, None))
        if len(_packed_cache) > 8:
            _packed_cache.clear()
        _packed_cache[key] = (ref, texture_arg._version, tex.data_ptr(), tex4)
    except TypeError:
        pass
    return tex4


def ensure_packed_texture(texture: torch.Tensor) -> Optional[torch.Tensor]:
    """Build (or reuse) the packed copy of ``texture`` on the CURRENT stream. Callers that are about to render the same
    texture from several CUDA streams do this before they fork, so that no stream reads a copy another one is still
    writing (texture_gs_b200.dist.render_views_accumulate(..., streams=n))."""
    if not USE_PACKED_TEXTURE or texture is None or not texture.is_cuda:
        return None
    lib = L.load()
    tex = _prep(texture, texture.device)
    with torch.cuda.device(texture.device):
        return _packed_texture(lib, texture, tex, torch.cuda.current_stream(texture.device).cuda_stream)


def packed_is_current(texture: torch.Tensor) -> bool:
    """True when the cached packed copy belongs to ``texture``'s current version (a forward would not repack)."""
    hit = _packed_cache.get(id(texture))
    return bool(hit is not None and hit[0]() is texture and hit[1] == texture._version and hit[2] == texture.data_ptr())


def packed_texture_buffer(texture: torch.Tensor) -> torch.Tensor:
    """A (6,R,R,4) buffer for the packed copy of ``texture`` — the cached one if there is one (its content is about
    to be replaced by the caller, see ``adopt_packed_texture``), else a new one."""
    hit = _packed_cache.get(id(texture))
    if hit is not None and hit[0]() is texture and hit[2] == texture.data_ptr() and hit[3].device == texture.device:
        return hit[3]
    return torch.empty(texture.shape[0], texture.shape[1], texture.shape[2], 4, device=texture.device, dtype=torch.float32)


def adopt_packed_texture(texture: torch.Tensor, tex4: torch.Tensor) -> None:
    """Register ``tex4`` as the packed copy of ``texture`` at its CURRENT version (texture_gs_b200.optim.TextureAdam
    writes it in the same kernel as the parameter update, so the next forward skips the repack)."""
    key = id(texture)
    try:
        ref = weakref.ref(texture, lambda _r, k=key: _packed_cache.pop(k, None))
    except TypeError:
        return
    if len(_packed_cache) > 8 and key not in _packed_cache:
        _packed_cache.clear()
    _packed_cache[key] = (ref, texture._version, texture.data_ptr(), tex4)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _prep(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.device != device:
        raise L.TexgsError(f"tensor on {t.device}, expected {device}")
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class RasterStats(NamedTuple):
    num_pairs: int
    num_visible: int
    max_tile_len: int
    num_blend: int
    pair_capacity: int


_last_stats = threading.local()
_spec = threading.local()


class SpecSwitches(NamedTuple):
    """The conventions of the path that the reference tree does not pin (its rasterizer source is not in
    the tree). Defaults = what the library computes unless told otherwise; each alternative is a cold instantiation of
    the render kernels with an oracle parity test (tests/test_gpu_parity.py), so the library can be re-pointed the day
    the upstream source says which one it is."""
    seamless_cube: bool = False           # E11-alt: taps beyond a face edge from the adjacent face (nvdiffrast 'cube' / GL seamless)
    depth_of_intersection: bool = False   # E7-alt: depth output = z of the ray-disc intersection, not of the centre
    stopgrad_delta: bool = False          # E13-alt: no gradient through the intersection offset
    upstream_clamp_grad: bool = False     # E2-alt: no gradient through a clamped x/z, y/z of the EWA projection (3DGS lineage)

    def flags(self) -> int:
        return ((L.FLAG_SEAMLESS_CUBE if self.seamless_cube else 0) | (L.FLAG_DEPTH_INTERSECTION if self.depth_of_intersection else 0)
                | (L.FLAG_STOPGRAD_DELTA if self.stopgrad_delta else 0) | (L.FLAG_CLAMP_GRAD_3DGS if self.upstream_clamp_grad else 0))


@contextmanager
def spec_switches(**kw):
    """``with spec_switches(seamless_cube=True): uv_tex_render(...)`` — rasterizer calls of this thread made inside the
    block use the given conventions (forward and its backward)."""
    prev = getattr(_spec, "v", SpecSwitches())
    _spec.v = prev._replace(**kw)
    try:
        yield _spec.v
    finally:
        _spec.v = prev


def current_spec() -> SpecSwitches:
    return getattr(_spec, "v", SpecSwitches())


_render_wait = threading.local()


@contextmanager
def wait_before_render(event: Optional["torch.cuda.Event"]):
    """Forward calls of this thread made inside the block wait for ``event`` between binning and the render kernel
    (``TexgsFwdArgs.render_wait_event``): preprocess, scan, scatter and sort of a view run while the optimizer step /
    gradient reduction of the previous batch is still in flight on another stream; the render — the first kernel that
    reads the texture — and the backward behind it are ordered after it."""
    prev = getattr(_render_wait, "v", None)
    _render_wait.v = event
    try:
        yield
    finally:
        _render_wait.v = prev


def last_stats() -> Optional[RasterStats]:
    """Counters (V, K, longest tile list, [debug] blended contributions) of the calling thread's
    most recent forward — what bench.py needs for the algorithmic-byte count."""
    return getattr(_last_stats, "v", None)


def _build_args(st: GaussianRasterizationSettings, mode: int, means3D, shs, colors_precomp, opacities, scales,
                rotations, uvs, gradient_uvs, texture, profile_arr=None, extra_attrs=None, cov3Ds_precomp=None, spec_flags: int = 0) -> L.TexgsFwdArgs:
    a = L.TexgsFwdArgs()
    a.P = means3D.shape[0]
    a.M = 0 if shs is None else shs.shape[1]
    a.sh_degree = int(st.sh_degree)
    a.E = 0 if extra_attrs is None else extra_attrs.shape[1]
    a.H, a.W = int(st.image_height), int(st.image_width)
    a.R = 0 if texture is None else texture.shape[1]
    a.mode = mode
    a.flags = (L.FLAG_PREFILTERED if st.prefiltered else 0) | (L.FLAG_DEBUG if st.debug else 0) | (spec_flags if mode == L.MODE_TEXTURE else (spec_flags & L.FLAG_CLAMP_GRAD_3DGS))
    a.tanfovx, a.tanfovy, a.scale_modifier = float(st.tanfovx), float(st.tanfovy), float(st.scale_modifier)
    a.viewmatrix = (C.c_float * 16)(*_host_floats(st.viewmatrix, 16))
    a.projmatrix = (C.c_float * 16)(*_host_floats(st.projmatrix, 16))
    a.campos = (C.c_float * 3)(*_host_floats(st.campos, 3))
    a.bg = (C.c_float * 3)(*_host_floats(st.bg, 3))
    a.means3D = _ptr(means3D)
    a.shs = _ptr(shs)
    a.colors_precomp = _ptr(colors_precomp)
    a.opacities = _ptr(opacities)
    a.scales = _ptr(scales)
    a.rotations = _ptr(rotations)
    a.uvs = _ptr(uvs)
    a.gradient_uvs = _ptr(gradient_uvs)
    a.texture = _ptr(texture)
    a.extra_attrs = _ptr(extra_attrs)
    a.cov3Ds_precomp = _ptr(cov3Ds_precomp)
    if profile_arr is not None:
        a.profile_events = C.cast(profile_arr, C.POINTER(C.c_void_p))
    ev = getattr(_render_wait, "v", None)
    if ev is not None:
        a.render_wait_event = C.c_void_p(ev.cuda_event)
    return a


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, uvs, gradient_uvs,
                texture, st: GaussianRasterizationSettings, mode: int, dual: bool = False, extra_attrs=None,
                cov3Ds_precomp=None):
        lib = L.load()
        if not means3D.is_cuda:
            raise L.TexgsError("the rasterizer runs on CUDA tensors only (no CPU fallback); got " + str(means3D.device))
        dev = means3D.device
        m3, sh, cp, op = _prep(means3D, dev), _prep(shs, dev), _prep(colors_precomp, dev), _prep(opacities, dev)
        sc, ro, uv, guv, tex = _prep(scales, dev), _prep(rotations, dev), _prep(uvs, dev), _prep(gradient_uvs, dev), _prep(texture, dev)
        P, H, W = m3.shape[0], int(st.image_height), int(st.image_width)
        if mode == L.MODE_TEXTURE:
            if tex is None or uv is None or guv is None:
                raise L.TexgsError("textured rasterization needs uvs, gradient_uvs and texture")
            if tex.dim() != 4 or tex.shape[0] != 6 or tex.shape[1] != tex.shape[2] or tex.shape[3] != 3:
                raise L.TexgsError(f"texture must be (6,R,R,3), got {tuple(tex.shape)}")
            if guv.shape != (P, 9) or uv.shape != (P, 3):
                raise L.TexgsError("uvs must be (P,3) and gradient_uvs (P,9)")
        if sh is not None and (sh.dim() != 3 or sh.shape[0] != P or sh.shape[2] != 3):
            raise L.TexgsError(f"shs must be (P,M,3), got {tuple(sh.shape)}")
        cov = _prep(cov3Ds_precomp, dev)
        if cov is not None:
            if sc is not None or ro is not None or mode == L.MODE_TEXTURE:
                raise L.TexgsError("cov3Ds_precomp replaces scales + rotations and is a diff_gauss (untextured) argument")
            if cov.shape != (P, 6):
                raise L.TexgsError(f"cov3Ds_precomp must be (P,6) [xx,xy,xz,yy,yz,zz], got {tuple(cov.shape)}")
        elif sc is None or ro is None or sc.shape != (P, 3) or ro.shape != (P, 4):
            raise L.TexgsError("scales (P,3) and rotations (P,4) expected")
        if op.numel() != P:
            raise L.TexgsError("opacities (P,1) expected")
        ex = _prep(extra_attrs, dev)
        if ex is not None and (ex.dim() != 2 or ex.shape[0] != P or ex.shape[1] < 1):
            raise L.TexgsError(f"extra_attrs must be (P,E) with E >= 1, got {tuple(ex.shape)}")

        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            from .profiling import current_event_array
            prof = current_event_array()      # captured here: backward runs on autograd's thread
            spec_flags = current_spec().flags()      # captured here: backward runs on autograd's thread
            a = _build_args(st, mode, m3, sh, cp, op, sc, ro, uv, guv, tex, prof, ex, cov, spec_flags)
            tex4 = None
            if mode == L.MODE_TEXTURE and USE_PACKED_TEXTURE:
                tex4 = _packed_texture(lib, texture, tex, stream)
                a.texture_rgba = _ptr(tex4)
            image = torch.empty(3, H, W, device=dev, dtype=torch.float32)
            depth = torch.empty(1, H, W, device=dev, dtype=torch.float32)
            norm = torch.empty(3, H, W, device=dev, dtype=torch.float32)
            alpha = torch.empty(1, H, W, device=dev, dtype=torch.float32)
            radii = torch.empty(P, device=dev, dtype=torch.int32)
            extra = torch.empty(ex.shape[1], H, W, device=dev, dtype=torch.float32) if ex is not None else image.new_empty(0)
            dual = bool(dual and mode == L.MODE_TEXTURE)
            image_nosh = torch.empty(3, H, W, device=dev, dtype=torch.float32) if dual else image.new_empty(0)
            if dual:
                a.out_image_nosh = _ptr(image_nosh)
            key = (dev.index, P, H, W)
            cap = _capacity_hint.get(key, max(1 << 16, 8 * P))
            pinned, event = ent = _acquire_counters(dev.index)
            gs, bs, is_ = C.c_size_t(), C.c_size_t(), C.c_size_t()
            for _attempt in range(8):
                L.check(lib.texgs_workspace_sizes(C.byref(a), cap, C.byref(gs), C.byref(bs), C.byref(is_)), "texgs_workspace_sizes")
                geom = torch.empty(max(gs.value, 256), device=dev, dtype=torch.uint8)
                binw = torch.empty(max(bs.value, 256), device=dev, dtype=torch.uint8)
                imgw = torch.empty(max(is_.value, 256), device=dev, dtype=torch.uint8)
                try:
                    L.check(lib.texgs_forward(C.byref(a), _ptr(geom), _ptr(binw), cap, _ptr(imgw), _ptr(image), _ptr(depth),
                                              _ptr(norm), _ptr(alpha), _ptr(radii), _ptr(extra) if ex is not None else None,
                                              C.c_void_p(pinned.data_ptr()), C.c_void_p(event.cuda_event), C.c_void_p(stream)),
                            "texgs_forward")
                    # waits only until the tile scan is done (early in the stream), not for the render
                    event.synchronize()
                    K, V, overflow, maxlen, blo, bhi = (int(x) & 0xFFFFFFFF for x in pinned[:6].tolist())
                except BaseException:
                    ent = None           # state of the pair unknown: drop it instead of returning it to the free list
                    raise
                if not overflow:
                    break
                cap = int(K * 1.25) + 4096
            else:
                raise L.TexgsError("pair capacity kept overflowing")
            if ent is not None:
                _release_counters(dev.index, ent)
            _capacity_hint[key] = max(cap, _capacity_hint.get(key, 0)) if not overflow else cap
            _last_stats.v = RasterStats(K, V, maxlen, blo | (bhi << 32), cap)

        ctx.st, ctx.mode, ctx.cap, ctx.prof, ctx.tex4, ctx.dual, ctx.spec_flags = st, mode, cap, prof, tex4, dual, spec_flags
        # fused gradient accumulation (texture_gs_b200.dist.GradBucket.fused): resolved now because
        # backward runs on autograd's thread. Only inputs that ARE bucket leaves qualify.
        ctx.fuse = None
        if any(ctx.needs_input_grad):
            from .dist import current_fused_bucket
            bucket = current_fused_bucket()
            if bucket is not None:
                ctx.fuse = {}
                for name, t in (("means3D", means3D), ("shs", shs), ("colors_precomp", colors_precomp), ("opacities", opacities),
                                ("scales", scales), ("rotations", rotations), ("uvs", uvs), ("texture", texture)):
                    if t is not None and t.is_leaf and t.requires_grad and t.is_contiguous() and t.dtype == torch.float32:
                        tgt = bucket.storage_for(t)
                        if tgt is not None and (name != "texture" or tgt[1] == (tex4 is not None)):
                            ctx.fuse[name] = tgt[0]
        ctx.dev = dev
        ctx.has = (shs is not None, colors_precomp is not None, uvs is not None, texture is not None)
        ctx.save_for_backward(m3, sh, cp, op, sc, ro, uv, guv, tex, ex, cov, geom, binw, imgw)
        nondiff = [radii]
        if not dual:
            nondiff.append(image_nosh)
        if ex is None:
            nondiff.append(extra)
        ctx.mark_non_differentiable(*nondiff)
        return image, depth, norm, alpha, radii, image_nosh, extra

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_image, g_depth, g_norm, g_alpha, _g_radii, g_image_nosh, g_extra):
        lib = L.load()
        m3, sh, cp, op, sc, ro, uv, guv, tex, ex, cov, geom, binw, imgw = ctx.saved_tensors
        st, mode, dev = ctx.st, ctx.mode, ctx.dev
        P = m3.shape[0]
        need = ctx.needs_input_grad   # means3D, means2D, shs, colors_precomp, opacities, scales, rotations, uvs, gradient_uvs, texture, st, mode, dual, extra_attrs, cov3Ds_precomp
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            b = L.TexgsBwdArgs()
            b.fwd = _build_args(st, mode, m3, sh, cp, op, sc, ro, uv, guv, tex, ctx.prof, ex, cov, ctx.spec_flags)
            if ctx.tex4 is not None:
                b.fwd.texture_rgba = _ptr(ctx.tex4)
            b.geom_ws, b.bin_ws, b.img_ws, b.pair_capacity = _ptr(geom), _ptr(binw), _ptr(imgw), ctx.cap
            keep = [_prep(g, dev) for g in (g_image, g_depth, g_norm, g_alpha)]
            b.dL_dimage, b.dL_ddepth, b.dL_dnorm, b.dL_dalpha = (_ptr(k) for k in keep)
            if ctx.dual:
                # the kernels only need the pointer STATE of the dual output to pick the variant
                b.fwd.out_image_nosh = C.c_void_p(1 << 8)
                keep.append(_prep(g_image_nosh, dev))
                b.dL_dimage_nosh = _ptr(keep[-1])
            d_ex = None
            if ex is not None:
                keep.append(_prep(g_extra, dev))
                b.dL_dextra = _ptr(keep[-1])
                if need[13]:
                    d_ex = torch.empty_like(ex)       # cleared by the library
                    b.dL_dextra_attrs = _ptr(d_ex)
            acc = torch.empty(max(P, 1) * L.BWD_ACC_FLOATS, device=dev, dtype=torch.float32)
            b.acc_ws = _ptr(acc)

            def out(flag, *shape):
                return torch.empty(*shape, device=dev, dtype=torch.float32) if flag else None

            fuse = ctx.fuse or {}
            accmask = 0

            def out_or_fused(name, bit, flag, *shape):
                # returns (tensor the kernel writes, gradient handed to autograd)
                nonlocal accmask
                if not flag:
                    return None, None
                if name in fuse:
                    accmask |= bit
                    return fuse[name], None
                t = torch.empty(*shape, device=dev, dtype=torch.float32)
                return t, t

            k_m3, d_m3 = out_or_fused("means3D", L.ACC_MEANS3D, need[0], P, 3)
            d_m2 = out(need[1], P, 3)
            k_sh, d_sh = out_or_fused("shs", L.ACC_SHS, need[2] and sh is not None, *(sh.shape if sh is not None else (0,)))
            k_cp, d_cp = out_or_fused("colors_precomp", L.ACC_COLORS, need[3] and cp is not None, P, 3)
            k_op, d_op = out_or_fused("opacities", L.ACC_OPACITY, need[4], P, 1)
            k_sc, d_sc = out_or_fused("scales", L.ACC_SCALES, need[5] and sc is not None, P, 3)
            k_ro, d_ro = out_or_fused("rotations", L.ACC_ROTATIONS, need[6] and ro is not None, P, 4)
            d_cov = out(need[14] and cov is not None, P, 6)
            b.dL_dcov3Ds = _ptr(d_cov)
            k_uv, d_uv = out_or_fused("uvs", L.ACC_UVS, need[7] and uv is not None, P, 3)
            d_tex, d_tex4 = None, None
            zero_tex = 1
            if need[9] and tex is not None and "texture" in fuse:
                zero_tex = 0                       # accumulate into the bucket; autograd gets no texture grad
                if ctx.tex4 is not None:
                    d_tex4 = fuse["texture"]
                else:
                    b.dL_dtexture = _ptr(fuse["texture"])
            elif need[9] and tex is not None:
                if ctx.tex4 is not None:
                    # padded gradient: 128-bit vector atomics in the kernel; the (6,R,R,3) gradient
                    # autograd sees is a strided view of it (no unpack pass)
                    d_tex4 = torch.empty(tex.shape[0], tex.shape[1], tex.shape[2], 4, device=dev, dtype=torch.float32)
                    d_tex = d_tex4[..., :3]
                else:
                    d_tex = torch.empty(tex.shape, device=dev, dtype=torch.float32)
            b.dL_dmeans3D, b.dL_dmeans2D, b.dL_dshs, b.dL_dcolors_precomp = _ptr(k_m3), _ptr(d_m2), _ptr(k_sh), _ptr(k_cp)
            b.dL_dopacity, b.dL_dscales, b.dL_drotations, b.dL_duvs = _ptr(k_op), _ptr(k_sc), _ptr(k_ro), _ptr(k_uv)
            if d_tex4 is not None:
                b.dL_dtexture_rgba = _ptr(d_tex4)
            elif d_tex is not None:
                b.dL_dtexture = _ptr(d_tex)
            b.zero_texture_grad = zero_tex
            b.accumulate_mask = accmask
            if zero_tex == 0:
                d_tex = None
            L.check(lib.texgs_backward(C.byref(b), C.c_void_p(stream)), "texgs_backward")
        return d_m3, d_m2, d_sh, d_cp, d_op, d_sc, d_ro, d_uv, None, d_tex, None, None, None, d_ex, d_cov


class GaussianRasterizer(nn.Module):
    """Callable with the kwargs of reference ``render/uv_tex_render.py:56-66`` (``uvs``,
    ``gradient_uvs``, ``texture`` given -> textured mode) or ``render/render.py:75-84``
    (``colors_precomp`` / full ``shs`` -> plain 3DGS mode)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        lib = L.load()
        st = self.raster_settings
        pos = _prep(positions, positions.device)
        if not pos.is_cuda:
            raise L.TexgsError("markVisible needs CUDA tensors")
        present = torch.empty(pos.shape[0], device=pos.device, dtype=torch.int32)
        with torch.cuda.device(pos.device):
            vm = (C.c_float * 16)(*_host_floats(st.viewmatrix, 16))
            pm = (C.c_float * 16)(*_host_floats(st.projmatrix, 16))
            L.check(lib.texgs_mark_visible(pos.shape[0], _ptr(pos), vm, pm, _ptr(present),
                                           C.c_void_p(torch.cuda.current_stream(pos.device).cuda_stream)), "texgs_mark_visible")
        return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3Ds_precomp=None, uvs=None, gradient_uvs=None, texture=None, extra_attrs=None, dual_no_sh=False):
        """``dual_no_sh=True`` (textured mode) appends a 7th result: the image the same
        splats give with ``sh_degree = 0``, blended in the same pass."""
        st = self.raster_settings
        if ((scales is None or rotations is None) and cov3Ds_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3Ds_precomp is not None):
            raise ValueError("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if cov3Ds_precomp is not None and texture is not None:
            raise ValueError("cov3Ds_precomp is a diff_gauss argument (render/render.py:83); the textured rasterizer needs scales and rotations")
        if texture is not None:
            mode = L.MODE_TEXTURE
            if colors_precomp is not None:
                raise ValueError("texture and colors_precomp are mutually exclusive")
        else:
            if (shs is None) == (colors_precomp is None):
                raise ValueError("Please provide exactly one of either SHs or precomputed colors!")
            mode = L.MODE_SH if shs is not None else L.MODE_PRECOMP
        if dual_no_sh and mode != L.MODE_TEXTURE:
            raise ValueError("dual_no_sh needs the textured mode")
        image, depth, norm, alpha, radii, image_nosh, extra = _RasterizeGaussians.apply(
            means3D, means2D, shs, colors_precomp, opacities, scales, rotations, uvs, gradient_uvs, texture, st, mode,
            bool(dual_no_sh), extra_attrs, cov3Ds_precomp)
        if extra_attrs is None:
            extra = None
        if dual_no_sh:
            return image, depth, norm, alpha, radii, extra, image_nosh
        return image, depth, norm, alpha, radii, extra
