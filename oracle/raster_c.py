"""ORACLE (test infrastructure, NOT product code) — loader of ``oracle/raster_c.c``, the scalar C + OpenMP restatement
of the textured rasterizer (forward and hand-derived backward, spec E1-E13; *parity unpinned*, see the C file's header).

    build(dtype)                      gcc -O2 -fopenmp ... -DREAL=double|float -> oracle/_build/libraster_c_f64|f32.so
    rasterize(..., cotangents=None)   same inputs / outputs as oracle.raster_ref.rasterize for the textured mode

Only ``tests/``, ``__graft_entry__`` (build + smoke) and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path
from typing import Optional, Sequence

import torch

HERE = Path(__file__).resolve().parent
SRC = HERE / "raster_c.c"
BUILD = HERE / "_build"
FLAG_THRESHOLD, FLAG_GRAZING, FLAG_FACE_TIE, FLAG_TEXEL_TIE, FLAG_DEPTH_TIE, FLAG_RECT, FLAG_CLAMP_TIE = 1, 2, 4, 8, 16, 32, 64
_libs = {}


def _args_struct(creal):
    p = C.POINTER(creal)

    class OracleArgs(C.Structure):
        _fields_ = [("P", C.c_int32), ("M", C.c_int32), ("sh_degree", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("R", C.c_int32),
                    ("threads", C.c_int32), ("reserved", C.c_int32),
                    ("tanfovx", creal), ("tanfovy", creal), ("scale_modifier", creal),
                    ("view", creal * 16), ("proj", creal * 16), ("campos", creal * 3), ("bg", creal * 3),
                    *[(n, p) for n in ("xyz", "shs", "opacity", "scaling", "rotation", "uvs", "grad_uvs", "texture")],
                    *[(n, p) for n in ("image", "depth", "norm", "alpha")],
                    ("radii", C.POINTER(C.c_int32)), ("final_T", p), ("n_contrib", C.POINTER(C.c_int32)), ("flags", C.POINTER(C.c_uint8)),
                    ("counters", C.POINTER(C.c_int64)),
                    *[(n, p) for n in ("g_image", "g_depth", "g_norm", "g_alpha")],
                    *[(n, p) for n in ("d_xyz", "d_means2D", "d_shs", "d_opacity", "d_scaling", "d_rotation", "d_uvs", "d_texture")]]
    return OracleArgs


def build(dtype=torch.float64, force: bool = False):
    """Compile (if the source is newer) and load the library for ``dtype`` (float64: the checker; float32: timing and
    fp32-arithmetic comparisons). Returns (lib, ArgsStruct, ctypes real type)."""
    f64 = dtype == torch.float64
    key = "f64" if f64 else "f32"
    if key in _libs and not force:
        return _libs[key]
    so = BUILD / f"libraster_c_{key}.so"
    if force or not so.exists() or so.stat().st_mtime < SRC.stat().st_mtime:
        gcc = shutil.which("gcc")
        if gcc is None:
            if not so.exists():
                raise RuntimeError("gcc not found and no prebuilt oracle library")
        else:
            BUILD.mkdir(exist_ok=True)
            tmp = so.with_suffix(f".{os.getpid()}.tmp")
            cmd = [gcc, "-O2", "-fopenmp", "-shared", "-fPIC", "-ffp-contract=off", "-std=c99", f"-DREAL={'double' if f64 else 'float'}",
                   str(SRC), "-o", str(tmp), "-lm"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + r.stderr[-4000:])
            os.replace(tmp, so)
    lib = C.CDLL(str(so))
    creal = C.c_double if f64 else C.c_float
    Args = _args_struct(creal)
    lib.texgs_oracle_run.restype = C.c_int
    lib.texgs_oracle_run.argtypes = [C.POINTER(Args)]
    lib.texgs_oracle_real_bytes.restype = C.c_int
    assert lib.texgs_oracle_real_bytes() == (8 if f64 else 4)
    _libs[key] = (lib, Args, creal)
    return _libs[key]


def rasterize(means3D, shs, opacities, scales, rotations, uvs, gradient_uvs, texture, settings, cotangents: Optional[Sequence] = None,
              dtype=torch.float64, threads: int = 0):
    """Textured render of ``settings`` (an ``oracle.raster_ref.RasterSettings`` / the reference's 12 fields). Returns
    ``(image, depth, norm, alpha, radii, aux)`` as CPU tensors of ``dtype``; ``aux`` holds final_T, n_contrib, the
    per-pixel conditioning flags (``ambiguous`` / ``grad_ambiguous`` with the meaning of oracle.raster_ref), counters
    and — when ``cotangents`` = (dL/dimage, dL/ddepth, dL/dnorm, dL/dalpha) is given — ``grads``."""
    lib, Args, creal = build(dtype)
    P = int(means3D.shape[0])
    H, W = int(settings.image_height), int(settings.image_width)
    Rr = int(texture.shape[1])
    keep = []

    def arr(t, shape=None):
        if t is None:
            return None, None
        x = t.detach().to(device="cpu", dtype=dtype).contiguous()
        if shape is not None:
            x = x.reshape(shape)
        keep.append(x)
        return x, C.cast(C.c_void_p(x.data_ptr()), C.POINTER(creal))

    a = Args()
    M = 0 if shs is None else int(shs.shape[1])
    a.P, a.M, a.sh_degree, a.H, a.W, a.R, a.threads = P, M, int(settings.sh_degree), H, W, Rr, int(threads)
    a.tanfovx, a.tanfovy, a.scale_modifier = float(settings.tanfovx), float(settings.tanfovy), float(settings.scale_modifier)
    a.view = (creal * 16)(*settings.viewmatrix.detach().double().reshape(-1).tolist())
    a.proj = (creal * 16)(*settings.projmatrix.detach().double().reshape(-1).tolist())
    a.campos = (creal * 3)(*settings.campos.detach().double().reshape(-1).tolist())
    a.bg = (creal * 3)(*settings.bg.detach().double().reshape(-1).tolist())
    for name, t in (("xyz", means3D), ("shs", shs), ("opacity", opacities), ("scaling", scales), ("rotation", rotations), ("uvs", uvs),
                    ("grad_uvs", gradient_uvs), ("texture", texture)):
        _, ptr = arr(t)
        if ptr is not None:
            setattr(a, name, ptr)
    out = {k: torch.zeros(s, dtype=dtype) for k, s in (("image", (3, H, W)), ("depth", (1, H, W)), ("norm", (3, H, W)), ("alpha", (1, H, W)),
                                                      ("final_T", (H, W)))}
    for k, v in out.items():
        setattr(a, k, C.cast(C.c_void_p(v.data_ptr()), C.POINTER(creal)))
    radii = torch.zeros(max(P, 1), dtype=torch.int32)
    n_contrib = torch.zeros(H, W, dtype=torch.int32)
    flags = torch.zeros(H, W, dtype=torch.uint8)
    counters = torch.zeros(4, dtype=torch.int64)
    a.radii = C.cast(C.c_void_p(radii.data_ptr()), C.POINTER(C.c_int32))
    a.n_contrib = C.cast(C.c_void_p(n_contrib.data_ptr()), C.POINTER(C.c_int32))
    a.flags = C.cast(C.c_void_p(flags.data_ptr()), C.POINTER(C.c_uint8))
    a.counters = C.cast(C.c_void_p(counters.data_ptr()), C.POINTER(C.c_int64))
    grads = None
    if cotangents is not None:
        for name, t, shape in zip(("g_image", "g_depth", "g_norm", "g_alpha"), cotangents, ((3, H, W), (H, W), (3, H, W), (H, W))):
            _, ptr = arr(t, shape)
            if ptr is not None:
                setattr(a, name, ptr)
        grads = {"xyz": torch.zeros(P, 3, dtype=dtype), "means2D": torch.zeros(P, 3, dtype=dtype), "opacity": torch.zeros(P, 1, dtype=dtype),
                 "scaling": torch.zeros(P, 3, dtype=dtype), "rotation": torch.zeros(P, 4, dtype=dtype), "uvs": torch.zeros(P, 3, dtype=dtype),
                 "texture": torch.zeros(6, Rr, Rr, 3, dtype=dtype), "shs": torch.zeros(P, M, 3, dtype=dtype) if M else None}
        for k, field in (("xyz", "d_xyz"), ("means2D", "d_means2D"), ("shs", "d_shs"), ("opacity", "d_opacity"), ("scaling", "d_scaling"),
                         ("rotation", "d_rotation"), ("uvs", "d_uvs"), ("texture", "d_texture")):
            if grads[k] is not None:
                setattr(a, field, C.cast(C.c_void_p(grads[k].data_ptr()), C.POINTER(creal)))
    rc = lib.texgs_oracle_run(C.byref(a))
    if rc != 0:
        raise RuntimeError(f"texgs_oracle_run failed ({rc})")
    fl = flags.numpy()
    value_flags = FLAG_THRESHOLD | FLAG_GRAZING | FLAG_FACE_TIE | FLAG_DEPTH_TIE | FLAG_RECT
    aux = dict(final_T=out["final_T"], n_contrib=n_contrib, flags=flags,
               ambiguous=torch.from_numpy((fl & value_flags) != 0), grad_ambiguous=torch.from_numpy(fl != 0),
               grazing=torch.from_numpy((fl & FLAG_GRAZING) != 0), texel_boundary=torch.from_numpy((fl & FLAG_TEXEL_TIE) != 0), rect_tie=torch.from_numpy((fl & FLAG_RECT) != 0), clamp_tie=torch.from_numpy((fl & FLAG_CLAMP_TIE) != 0),
               num_pairs=int(counters[0]), num_visible=int(counters[1]), num_blend=int(counters[2]), max_tile_len=int(counters[3]),
               grads=grads)
    return out["image"], out["depth"], out["norm"], out["alpha"], radii[:P], aux
