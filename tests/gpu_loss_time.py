"""Kernel-only timing of the fused photometric loss at 1080p through the C-ABI (TEXGS_LIB selects the build); one JSON line."""
import ctypes as C, json, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from texture_gs_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda")
Cc, H, W = 3, 1080, 1920
img, gt = torch.rand(Cc, H, W, device=dev), torch.rand(Cc, H, W, device=dev)
nb = C.c_size_t()
lib.texgs_photometric_workspace_size(Cc, H, W, C.byref(nb))
ws = torch.empty(nb.value, device=dev, dtype=torch.uint8)
out3, coef, dimg = torch.empty(3, device=dev), torch.tensor([0.8, 0.2], device=dev), torch.empty_like(img)
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
fwd = lambda: lib.texgs_photometric_forward(p(img), p(gt), Cc, H, W, C.c_float(0.2), p(ws), p(out3), st)
bwd = lambda: lib.texgs_photometric_backward(p(img), p(gt), Cc, H, W, p(ws), p(coef), p(dimg), st)
res = {"lib": os.environ.get("TEXGS_LIB", "default")}
for name, f in (("forward_ms", fwd), ("backward_ms", bwd)):
    for _ in range(5):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        f()
    e1.record(); torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / 50, 4)
res["loss"] = out3.tolist(); res["dimg_abs_sum"] = float(dimg.abs().sum())
print(json.dumps(res))
