"""Kernel-only timing of texgs_texture_adam_step at R = 2048 (TEXGS_LIB selects the build); one JSON line.
Bytes per texel: p, m, v read+write 72, gradient read 16 (padded) or 12, + 16 zeroing, + 16 packed copy."""
import ctypes as C, json, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from texture_gs_b200 import _lib as L
lib = L.load()
R = 2048
n = 6 * R * R
p, m, v = (torch.randn(n * 3, device="cuda") for _ in range(3))
v.abs_()
g4, rgba, g3 = torch.randn(n * 4, device="cuda"), torch.empty(n * 4, device="cuda"), torch.randn(n * 3, device="cuda")
ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = {"lib": os.environ.get("TEXGS_LIB", "default")}
for name, (G3, G4, RG, Z, nbytes) in {"padded_zero_rgba": (None, g4, rgba, 1, 120), "padded_only": (None, g4, None, 0, 88),
                                      "plain_grad_only": (g3, None, None, 0, 84)}.items():
    def step(i):
        L.check(lib.texgs_texture_adam_step(ptr(p), ptr(m), ptr(v), ptr(G3), ptr(G4), ptr(RG), n, 0.0025, 0.9, 0.999, 1e-15, i + 1, Z, stream), "adam")
    for i in range(3): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): step(i + 3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[name] = {"ms": round(ms, 4), "GBps": round(n * nbytes / ms / 1e6, 0)}
print(json.dumps(out))
