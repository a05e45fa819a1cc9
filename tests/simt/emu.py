"""TEST INFRASTRUCTURE: builds the CUDA sources of libtexgs for the HOST on top of tests/simt/simt_emu.h and drives
the resulting library (same C-ABI, host pointers) so that CPU-only tests can run the real kernel source.

* ``build()`` rewrites ``texgs_api.cu`` textually (``k<<<g, b, s, st>>>(args)`` -> ``SIMT_LAUNCH(k, g, b, s, st, args)``,
  the tcgen05 UV-MLP section dropped: tensor-core PTX cannot be emulated), compiles it with g++ into
  ``tests/simt/_build/libtexgs_emu_<hash>.so`` and loads it with the argument types of ``texture_gs_b200._lib``.
* ``rasterize(...)`` mirrors the forward/backward call sequence of ``texture_gs_b200/rasterizer.py`` with CPU tensors.

Only tests import this module; the product never does (it raises without a CUDA device)."""
from __future__ import annotations

import ctypes as C
import hashlib
import shutil
import subprocess
from pathlib import Path
from typing import Optional

import torch

from texture_gs_b200 import _lib as L

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "texture_gs_b200" / "csrc"
BUILD = HERE / "_build"
DYN_SMEM = (("unsigned char", "smem_raw"), ("float", "prebwd_smem"), ("float", "prefwd_smem"))     # every `extern __shared__` array of the sources
GXX_FLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes",
             "-fno-extern-tls-init",     # `extern __shared__` arrays are plain extern thread_local arrays: no init wrapper
             "-fno-gnu-unique", "-Wl,-Bsymbolic"]   # two builds (different -D flags) loaded into one process keep their own state
SKIPPED_SYMBOLS = ("texgs_uvmlp_forward", "texgs_uvmlp_backward_head", "texgs_uvmlp_backward_layer", "texgs_uvmlp_backward_tail")


def _split_top_level(s: str):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(src: str) -> str:
    """kernel<T...><<<grid, block, smem, stream>>>(args...)  ->  SIMT_LAUNCH((kernel<T...>), grid, block, smem, stream, args...)"""
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            return out + src[pos:]
        # kernel expression: identifier, optionally followed by one template argument list
        k = i
        if src[k - 1] == ">":
            depth = 0
            while True:
                k -= 1
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        kernel = src[k:i]
        j = src.index(">>>", i)
        cfg = _split_top_level(src[i + 3:j])
        assert len(cfg) == 4, ("launch without an explicit <<<grid, block, smem, stream>>>", src[k:j + 3])
        a = j + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", src[k:a + 1]
        depth, e = 0, a
        while True:
            if src[e] == "(":
                depth += 1
            elif src[e] == ")":
                depth -= 1
                if depth == 0:
                    break
            e += 1
        args = src[a + 1:e].strip()
        out += src[pos:k] + f"SIMT_LAUNCH(({kernel}), {cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}" + (", " + args if args else "") + ")"
        pos = e + 1


def emulated_source() -> str:
    api = (CSRC / "texgs_api.cu").read_text()
    api = api.replace('#include "texgs_uvmlp.cuh"\n', "")
    a = api.index("int texgs_uvmlp_forward(")
    b = api.index("__global__ void texgs_mark_visible_kernel")
    api = api[:a] + api[b:]
    api = rewrite_launches(api)
    head = '#include "simt_emu.h"\n' + "".join(f"SIMT_DEFINE_DYN_SMEM({t}, {n})\n" for t, n in DYN_SMEM)
    tail = ('\nextern "C" const char* simt_last_error(void) { return simt::G().err_msg.c_str(); }\n'
            'extern "C" unsigned long long simt_collectives(void) { return simt::G().collectives; }\n'
            'extern "C" void simt_set_eager_copies(int on) { simt::G().eager_copies = on != 0; }\n'
            'extern "C" void simt_profile_votes(int on) { simt::G().profile_votes = on != 0; if (on) { simt::G().votes.clear(); simt::G().ballot_trace.clear(); } }\n'
            'extern "C" unsigned long long simt_ballot_trace(unsigned* out, unsigned long long cap) { auto& t = simt::G().ballot_trace;\n'
            '    if (out) for (size_t i = 0; i < t.size() && i < cap; ++i) out[i] = t[i]; return t.size(); }\n'
            'extern "C" int simt_vote_sites(void) { return (int)simt::G().votes.size(); }\n'
            'extern "C" void simt_vote_dump(unsigned long long* sites, unsigned long long* hist34) {\n'
            '    size_t i = 0; for (auto& kv : simt::G().votes) { sites[i] = (unsigned long long)(uintptr_t)kv.first; hist34[34 * i] = kv.second.calls;\n'
            '        for (int k = 0; k < 33; ++k) hist34[34 * i + 1 + k] = kv.second.hist[k]; ++i; } }\n'
            'extern "C" const void* simt_anchor(void) { return (const void*)&simt_anchor; }\n'
            'extern "C" void simt_set_schedule_seed(unsigned seed) { simt::G().sched_seed = seed; simt::G().sched_state = seed * 2654435761u + 1u; }\n'
            'extern "C" void simt_set_fastmath_noise(unsigned ulps) { simt::G().fastmath_noise_ulps = ulps; simt::G().noise_state = 0x9e3779b9u; }\n')
    return head + api + tail


_loaded = {}


def build(extra_flags=()):
    """Compile (cached on the hash of every input) and load the emulated library."""
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    src = emulated_source()
    h = hashlib.sha256()
    h.update(src.encode())
    h.update((HERE / "simt_emu.h").read_bytes())
    for f in sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "texgs.h"]:
        h.update(f.read_bytes())
    h.update(" ".join([*GXX_FLAGS, *extra_flags]).encode())
    tag = h.hexdigest()[:16]
    if tag in _loaded:
        return _loaded[tag]
    BUILD.mkdir(exist_ok=True)
    so = BUILD / f"libtexgs_emu_{tag}.so"
    if not so.exists():
        cpp = BUILD / f"texgs_emu_{tag}.cpp"
        cpp.write_text(src)
        cmd = [gxx, *GXX_FLAGS, "-I", str(HERE), "-I", str(CSRC), *extra_flags, str(cpp), "-o", str(so)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stderr[-6000:])
        keep = {so, cpp} | {BUILD / f"libtexgs_emu_{t}.so" for t in _loaded} | {BUILD / f"texgs_emu_{t}.cpp" for t in _loaded}
        for old in list(BUILD.glob("libtexgs_emu_*.so")) + list(BUILD.glob("texgs_emu_*.cpp")):
            if old not in keep:
                old.unlink()
    lib = C.CDLL(str(so))
    for name, (res, args) in L.SYMBOLS.items():
        if name in SKIPPED_SYMBOLS:
            continue
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib.simt_last_error.restype = C.c_char_p
    lib.simt_collectives.restype = C.c_ulonglong
    lib.simt_set_eager_copies.argtypes = [C.c_int]
    lib.simt_set_schedule_seed.argtypes = [C.c_uint]
    lib.simt_profile_votes.argtypes = [C.c_int]
    lib.simt_vote_sites.restype = C.c_int
    lib.simt_vote_dump.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.simt_anchor.restype = C.c_void_p
    lib.simt_ballot_trace.argtypes = [C.POINTER(C.c_uint), C.c_ulonglong]
    lib.simt_ballot_trace.restype = C.c_ulonglong
    lib._so_path = str(so)
    lib.simt_set_fastmath_noise.argtypes = [C.c_uint]
    _loaded[tag] = lib
    return lib


def check(lib, rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib.texgs_last_error().decode('utf-8', 'replace')}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _aligned_empty(nbytes: int, align: int = 256) -> torch.Tensor:
    buf = torch.empty(nbytes + align, dtype=torch.uint8)
    off = (-buf.data_ptr()) % align
    return buf[off:off + nbytes]


def _f32(t, align: int = 16):
    """contiguous fp32 CPU copy whose storage is ``align``-byte aligned (the library checks 16-byte alignment)."""
    if t is None:
        return None
    t = t.detach().to(torch.float32).contiguous()
    if t.data_ptr() % align:
        raw = _aligned_empty(t.numel() * 4, align)
        out = raw.view(torch.float32).view(t.shape)
        out.copy_(t)
        return out
    return t


class EmuResult:
    pass


def rasterize(*, means3D, opacities, scales=None, rotations=None, shs=None, colors_precomp=None, uvs=None, gradient_uvs=None,
              texture=None, extra_attrs=None, cov3Ds_precomp=None, H, W, tanfovx, tanfovy, bg, scale_modifier=1.0,
              viewmatrix, projmatrix, campos, sh_degree, cotangents=None, dual_no_sh=False, packed_texture=True,
              packed_texture_grad=True, debug=False, cot_nosh=None, cot_extra=None, pair_capacity=None, lib=None,
              accumulate_onto=None, spec_flags=0):
    """Forward (+ backward when ``cotangents`` = (dL/dimage, dL/ddepth, dL/dnorm, dL/dalpha) is given) of the emulated
    library on CPU tensors; the call sequence is the one of texture_gs_b200/rasterizer.py.

    ``accumulate_onto``: a float; every gradient buffer the library can accumulate into (``accumulate_mask``, and the
    texture gradient with ``zero_texture_grad = 0``) is pre-filled with it — the fused-bucket path of GradBucket."""
    lib = lib or build()
    m3, op, sc, ro, sh, cp = _f32(means3D), _f32(opacities), _f32(scales), _f32(rotations), _f32(shs), _f32(colors_precomp)
    uv, guv, tex, ex, cov = _f32(uvs), _f32(gradient_uvs), _f32(texture), _f32(extra_attrs), _f32(cov3Ds_precomp)
    P = m3.shape[0]
    mode = L.MODE_TEXTURE if tex is not None else (L.MODE_SH if sh is not None else L.MODE_PRECOMP)
    a = L.TexgsFwdArgs()
    a.P, a.M, a.sh_degree, a.E = P, (0 if sh is None else sh.shape[1]), int(sh_degree), (0 if ex is None else ex.shape[1])
    a.H, a.W, a.R, a.mode = int(H), int(W), (0 if tex is None else tex.shape[1]), mode
    a.flags = (L.FLAG_DEBUG if debug else 0) | int(spec_flags)
    a.tanfovx, a.tanfovy, a.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
    a.viewmatrix = (C.c_float * 16)(*viewmatrix.reshape(-1).tolist())
    a.projmatrix = (C.c_float * 16)(*projmatrix.reshape(-1).tolist())
    a.campos = (C.c_float * 3)(*campos.reshape(-1).tolist())
    a.bg = (C.c_float * 3)(*[float(x) for x in bg])
    a.means3D, a.shs, a.colors_precomp, a.opacities, a.scales, a.rotations = _ptr(m3), _ptr(sh), _ptr(cp), _ptr(op), _ptr(sc), _ptr(ro)
    a.uvs, a.gradient_uvs, a.texture, a.extra_attrs, a.cov3Ds_precomp = _ptr(uv), _ptr(guv), _ptr(tex), _ptr(ex), _ptr(cov)
    tex4 = None
    if mode == L.MODE_TEXTURE and packed_texture:
        tex4 = _aligned_empty(6 * a.R * a.R * 16).view(torch.float32).view(6, a.R, a.R, 4)
        check(lib, lib.texgs_pack_texture(_ptr(tex), a.R, _ptr(tex4), None), "texgs_pack_texture")
        a.texture_rgba = _ptr(tex4)
    image, depth, norm, alpha = torch.empty(3, H, W), torch.empty(1, H, W), torch.empty(3, H, W), torch.empty(1, H, W)
    radii = torch.empty(max(P, 1), dtype=torch.int32)
    extra = torch.empty(ex.shape[1], H, W) if ex is not None else None
    dual = bool(dual_no_sh and mode == L.MODE_TEXTURE)
    image_nosh = torch.empty(3, H, W) if dual else None
    if dual:
        a.out_image_nosh = _ptr(image_nosh)
    cap = int(pair_capacity) if pair_capacity is not None else max(1 << 12, 8 * P)
    counters = L.TexgsCounters()
    gs, bs, is_ = C.c_size_t(), C.c_size_t(), C.c_size_t()
    for _ in range(8):
        check(lib, lib.texgs_workspace_sizes(C.byref(a), cap, C.byref(gs), C.byref(bs), C.byref(is_)), "texgs_workspace_sizes")
        geom, binw, imgw = _aligned_empty(max(gs.value, 256)), _aligned_empty(max(bs.value, 256)), _aligned_empty(max(is_.value, 256))
        geom.fill_(0xA5); binw.fill_(0xA5); imgw.fill_(0xA5)            # workspaces arrive uninitialised on the GPU too
        check(lib, lib.texgs_forward(C.byref(a), _ptr(geom), _ptr(binw), cap, _ptr(imgw), _ptr(image), _ptr(depth), _ptr(norm),
                                     _ptr(alpha), _ptr(radii), _ptr(extra), C.cast(C.pointer(counters), C.c_void_p), None, None),
              "texgs_forward")
        if not counters.overflow:
            break
        cap = int(counters.num_pairs * 1.25) + 64
    else:
        raise RuntimeError("pair capacity kept overflowing")
    res = EmuResult()
    res.image, res.depth, res.norm, res.alpha, res.radii, res.extra, res.image_nosh = image, depth, norm, alpha, radii[:P], extra, image_nosh
    res.num_pairs, res.num_visible, res.max_tile_len = counters.num_pairs, counters.num_visible, counters.max_tile_len
    res.num_blend = counters.num_blend_lo | (counters.num_blend_hi << 32)
    res.pair_capacity = cap
    lay = L.TexgsLayout()
    check(lib, lib.texgs_workspace_layout(C.byref(a), cap, C.byref(lay)), "texgs_workspace_layout")
    T = int(lay.num_tiles)
    res.tile_offset = binw[lay.bin_tile_offset:lay.bin_tile_offset + 4 * (T + 1)].view(torch.int32).clone()
    res.sorted_ids = binw[lay.bin_sorted_ids:lay.bin_sorted_ids + 4 * counters.num_pairs].view(torch.int32).clone()
    res.grads = None
    if cotangents is None:
        return res

    b = L.TexgsBwdArgs()
    b.fwd = a
    b.geom_ws, b.bin_ws, b.img_ws, b.pair_capacity = _ptr(geom), _ptr(binw), _ptr(imgw), cap
    keep = [_f32(c) for c in cotangents]
    b.dL_dimage, b.dL_ddepth, b.dL_dnorm, b.dL_dalpha = (_ptr(k) for k in keep)
    if dual:
        keep.append(_f32(cot_nosh if cot_nosh is not None else torch.zeros(3, H, W)))
        b.dL_dimage_nosh = _ptr(keep[-1])
    g = {}
    if ex is not None:
        keep.append(_f32(cot_extra if cot_extra is not None else torch.zeros_like(extra)))
        b.dL_dextra = _ptr(keep[-1])
        g["extra_attrs"] = _aligned_empty(ex.numel() * 4).view(torch.float32).view(ex.shape)
        b.dL_dextra_attrs = _ptr(g["extra_attrs"])
    acc = _aligned_empty(max(P, 1) * L.BWD_ACC_FLOATS * 4).view(torch.float32)
    acc.fill_(float("nan"))                                            # cleared by the library
    b.acc_ws = _ptr(acc)

    def out(name, *shape):
        n = 1
        for s in shape:
            n *= s
        g[name] = _aligned_empty(max(n, 1) * 4).view(torch.float32)[:n].view(*shape)
        g[name].fill_(float("nan") if (accumulate_onto is None or name == "means2D") else float(accumulate_onto))
        return _ptr(g[name])

    b.dL_dmeans3D, b.dL_dmeans2D, b.dL_dopacity = out("means3D", P, 3), out("means2D", P, 3), out("opacities", P, 1)
    if sc is not None:
        b.dL_dscales, b.dL_drotations = out("scales", P, 3), out("rotations", P, 4)
    if cov is not None:
        b.dL_dcov3Ds = out("cov3Ds_precomp", P, 6)
    if sh is not None:
        b.dL_dshs = out("shs", *sh.shape)
    if cp is not None:
        b.dL_dcolors_precomp = out("colors_precomp", P, 3)
    if mode == L.MODE_TEXTURE:
        b.dL_duvs = out("uvs", P, 3)
        if packed_texture_grad:
            b.dL_dtexture_rgba = out("texture_rgba", 6, a.R, a.R, 4)
        else:
            b.dL_dtexture = out("texture", 6, a.R, a.R, 3)
    b.zero_texture_grad, b.accumulate_mask = 1, 0
    if accumulate_onto is not None:
        b.zero_texture_grad = 0
        b.accumulate_mask = (L.ACC_MEANS3D | L.ACC_OPACITY | L.ACC_SCALES | L.ACC_ROTATIONS | L.ACC_SHS | L.ACC_COLORS | L.ACC_UVS)
    check(lib, lib.texgs_backward(C.byref(b), None), "texgs_backward")
    if "texture_rgba" in g:
        pad_expected = 0.0 if accumulate_onto is None else float(accumulate_onto)
        assert bool((g["texture_rgba"][..., 3] == pad_expected).all()), "padding channel of the texel gradient must stay untouched"
        g["texture"] = g.pop("texture_rgba")[..., :3].contiguous()
    res.grads = g
    return res


def vote_profile(lib):
    """After ``lib.simt_profile_votes(1)`` and some launches: {source location of every __ballot/__any/__all_sync call site:
    (calls, histogram[0..32] of the number of lanes that voted true)}. Locations come from addr2line on the emulated
    library (built with -g), i.e. they name lines of the .cuh kernels."""
    n = lib.simt_vote_sites()
    sites = (C.c_ulonglong * max(n, 1))()
    hist = (C.c_ulonglong * (34 * max(n, 1)))()
    lib.simt_vote_dump(sites, hist)
    # load base of the shared object: the anchor symbol's address minus its offset in the file
    off = int(subprocess.run(["nm", "-D", "--defined-only", lib._so_path], capture_output=True, text=True).stdout.split(" T simt_anchor")[0].split()[-1], 16)
    base = lib.simt_anchor() - off
    out = {}
    for i in range(n):
        addr = sites[i] - base - 1
        r = subprocess.run(["addr2line", "-e", lib._so_path, "-f", "-C", "-i", hex(addr)], capture_output=True, text=True).stdout.strip().splitlines()
        loc = next((ln for ln in r if ".cuh:" in ln or ".cu:" in ln), r[-1] if r else "?")
        fn = r[0] if r else "?"
        key = f"{Path(loc.split(' ')[0]).name} [{fn.split('(')[0]}]"
        calls = hist[34 * i]
        h = [hist[34 * i + 1 + k] for k in range(33)]
        if key in out:
            calls += out[key][0]
            h = [a + b for a, b in zip(h, out[key][1])]
        out[key] = (calls, h)
    return out


def ballot_trace(lib):
    """(n, 4) int64 array, columns (launch number, block, warp, lanes true), one row for every
    __ballot_sync executed since ``simt_profile_votes(1)``, in execution order."""
    import numpy as np
    n = int(lib.simt_ballot_trace(None, 0))
    buf = (C.c_uint * max(n, 1))()
    lib.simt_ballot_trace(buf, n)
    a = np.frombuffer(buf, dtype=np.uint32, count=n).reshape(-1, 3).astype(np.int64)
    return np.stack([a[:, 0], a[:, 1] >> 8, a[:, 1] & 255, a[:, 2]], axis=1)
