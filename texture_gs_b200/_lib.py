"""ctypes binding of libtexgs.so (C-ABI declared in include/texgs.h).

The product path has NO fallback: if the shared library is missing or a symbol is absent this
module raises, and every rasterizer call fails loudly (it never routes through ``oracle/``).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# TEXGS_LIB selects an alternative build of the same ABI (kernel-tuning experiments, tools/build_variants.py)
LIB_PATH = Path(os.environ["TEXGS_LIB"]).resolve() if os.environ.get("TEXGS_LIB") else _PKG / "libtexgs.so"

TEXGS_ABI_VERSION = 4
FLAG_PREFILTERED = 1
FLAG_DEBUG = 2
FLAG_SEAMLESS_CUBE = 4        # spec switches (include/texgs.h): E11-alt, E7-alt, E13-alt
FLAG_DEPTH_INTERSECTION = 8
FLAG_STOPGRAD_DELTA = 16
FLAG_CLAMP_GRAD_3DGS = 32
MODE_TEXTURE, MODE_SH, MODE_PRECOMP = 0, 1, 2
BWD_ACC_FLOATS = 24
ACC_MEANS3D, ACC_MEANS2D, ACC_OPACITY, ACC_SCALES, ACC_ROTATIONS, ACC_SHS, ACC_COLORS, ACC_UVS = 1, 2, 4, 8, 16, 32, 64, 128
EV_COUNT = 10
EV_NAMES = ("fwd_start", "preprocess_fwd", "scan_tiles", "scatter_pairs", "sort_tiles", "render_fwd",
            "bwd_start", "bwd_clear", "render_bwd", "preprocess_bwd")

_fp = C.c_void_p   # device pointers travel as void*


class TexgsFwdArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("M", C.c_int32), ("sh_degree", C.c_int32), ("E", C.c_int32),
        ("H", C.c_int32), ("W", C.c_int32), ("R", C.c_int32), ("mode", C.c_int32),
        ("flags", C.c_uint32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("viewmatrix", C.c_float * 16), ("projmatrix", C.c_float * 16),
        ("campos", C.c_float * 3), ("bg", C.c_float * 3),
        ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("opacities", _fp), ("scales", _fp),
        ("rotations", _fp), ("uvs", _fp), ("gradient_uvs", _fp), ("texture", _fp), ("extra_attrs", _fp),
        ("cov3Ds_precomp", _fp),
        ("texture_rgba", _fp),
        ("out_image_nosh", _fp),
        ("profile_events", C.POINTER(C.c_void_p)),
        ("render_wait_event", C.c_void_p),
    ]


class TexgsCounters(C.Structure):
    _fields_ = [("num_pairs", C.c_uint32), ("num_visible", C.c_uint32), ("overflow", C.c_uint32),
                ("max_tile_len", C.c_uint32), ("num_blend_lo", C.c_uint32), ("num_blend_hi", C.c_uint32),
                ("num_long_tiles", C.c_uint32), ("reserved", C.c_uint32 * 1)]


class TexgsBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", TexgsFwdArgs),
        ("geom_ws", _fp), ("bin_ws", _fp), ("img_ws", _fp), ("pair_capacity", C.c_uint64),
        ("dL_dimage", _fp), ("dL_ddepth", _fp), ("dL_dnorm", _fp), ("dL_dalpha", _fp), ("dL_dextra", _fp), ("dL_dimage_nosh", _fp),
        ("acc_ws", _fp),
        ("dL_dmeans3D", _fp), ("dL_dmeans2D", _fp), ("dL_dopacity", _fp), ("dL_dscales", _fp),
        ("dL_drotations", _fp), ("dL_dshs", _fp), ("dL_dcolors_precomp", _fp), ("dL_duvs", _fp),
        ("dL_dtexture", _fp), ("dL_dtexture_rgba", _fp), ("dL_dextra_attrs", _fp), ("dL_dcov3Ds", _fp),
        ("zero_texture_grad", C.c_int32), ("accumulate_mask", C.c_uint32),
    ]


class TexgsLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "geom_records", "geom_rects", "bin_counters", "bin_tile_count", "bin_tile_offset", "bin_tile_cursor",
        "bin_pairs", "bin_sorted_ids", "bin_cull_masks", "img_final_T", "img_n_contrib", "num_tiles", "record_bytes")]


class TexgsUvMlpArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("xyz", C.c_void_p), ("offset", C.c_float * 3), ("inv_scale", C.c_float * 3),
                ("W1", C.c_void_p), ("b1", C.c_void_p), ("W_hidden", C.c_void_p * 3), ("b_hidden", C.c_void_p * 3),
                ("emb", C.c_void_p), ("W5", C.c_void_p), ("b5", C.c_void_p), ("uv", C.c_void_p), ("jacobian", C.c_void_p),
                ("stash", C.c_void_p * 4), ("stash_inv_len", C.c_void_p), ("debug_accumulators", C.c_void_p)]


class TexgsDpAdamArgs(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("grad_ptrs", C.c_void_p * 16), ("param_ptrs", C.c_void_p * 16),
                ("grad_mc", C.c_void_p), ("param_mc", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n_texels", C.c_uint64), ("tile_lo", C.c_uint64), ("tile_hi", C.c_uint64),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("step", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/texgs.h declares: (name, restype, argtypes)
SYMBOLS = {
    "texgs_abi_version": (C.c_int, []),
    "texgs_last_error": (C.c_char_p, []),
    "texgs_kernel_names": (C.c_char_p, []),
    "texgs_workspace_sizes": (C.c_int, [C.POINTER(TexgsFwdArgs), C.c_uint64, C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "texgs_workspace_layout": (C.c_int, [C.POINTER(TexgsFwdArgs), C.c_uint64, C.POINTER(TexgsLayout)]),
    "texgs_forward": (C.c_int, [C.POINTER(TexgsFwdArgs), _fp, _fp, C.c_uint64, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "texgs_backward": (C.c_int, [C.POINTER(TexgsBwdArgs), C.c_void_p]),
    "texgs_pack_texture": (C.c_int, [_fp, C.c_int32, _fp, C.c_void_p]),
    "texgs_photometric_workspace_size": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "texgs_photometric_forward": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_float, _fp, _fp, C.c_void_p]),
    "texgs_photometric_backward": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, C.c_int32, _fp, _fp, _fp, C.c_void_p]),
    "texgs_geometry_loss_workspace_size": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "texgs_geometry_loss_forward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int32, C.c_int32, C.c_float, _fp, _fp, C.c_void_p]),
    "texgs_geometry_loss_backward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int32, C.c_int32, C.c_float, _fp, _fp, _fp, _fp, C.c_void_p]),
    "texgs_texture_adam_step": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_int32, C.c_int32, C.c_void_p]),
    "texgs_dp_shard": (C.c_int, [C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "texgs_texture_adam_dp_step": (C.c_int, [C.POINTER(TexgsDpAdamArgs), C.c_void_p]),
    "texgs_uvmlp_forward": (C.c_int, [C.POINTER(TexgsUvMlpArgs), C.c_void_p]),
    "texgs_uvmlp_backward_head": (C.c_int, [C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_void_p]),
    "texgs_uvmlp_backward_layer": (C.c_int, [C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, C.c_void_p]),
    "texgs_uvmlp_backward_tail": (C.c_int, [C.c_int32, _fp, _fp, C.POINTER(C.c_float), C.POINTER(C.c_float), _fp, _fp, _fp, _fp, C.c_void_p]),
    "texgs_mark_visible": (C.c_int, [C.c_int32, _fp, C.POINTER(C.c_float), C.POINTER(C.c_float), _fp, C.c_void_p]),
}

_lib = None


class TexgsError(RuntimeError):
    pass


def load():
    """Load libtexgs.so (once). Raises TexgsError if it is not built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TexgsError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `python -m texture_gs_b200.build`). The rasterizer has no CPU / PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise TexgsError(f"libtexgs.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    v = lib.texgs_abi_version()
    if v != TEXGS_ABI_VERSION:
        raise TexgsError(f"libtexgs.so ABI {v} != expected {TEXGS_ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().texgs_last_error().decode("utf-8", "replace")
        raise TexgsError(f"{what} failed (code {rc}): {msg}")
