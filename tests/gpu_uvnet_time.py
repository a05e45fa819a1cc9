"""Kernel-only timing of the fused UV + Jacobian producer at 500 k points (TEXGS_LIB selects the build); one JSON line."""
import json, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import uvnet_ref as UR
from texture_gs_b200.uvnet import FusedUVNet
n = 500_000
net = FusedUVNet(bias=False).cuda()
net.load_state_dict(UR.random_params(seed=0, bias=False))
xyz, emb = torch.randn(n, 3, device="cuda"), 0.5 * torch.randn(128, device="cuda")
out = {"lib": os.environ.get("TEXGS_LIB", "default"), "points": n}
for name, grad in (("inference", False), ("training_forward_with_stash", True)):
    x = xyz.clone().requires_grad_(grad)
    with torch.set_grad_enabled(grad):
        for _ in range(3):
            net.uv_and_jacobian(x, emb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            net.uv_and_jacobian(x, emb)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[name] = {"ms": round(ms, 4), "tensor_TFLOPs": round(n * 4 * 2 * (3 * 128 * 128 + 16 * 128) / ms / 1e9, 1)}
x = xyz.clone().requires_grad_(True)
cot = torch.randn(n, 3, device="cuda")
for _ in range(2):
    net.zero_grad(); (net(x, emb) * cot).sum().backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    net.zero_grad(); (net(x, emb) * cot).sum().backward()
e1.record(); torch.cuda.synchronize()
out["forward_plus_backward_ms"] = round(e0.elapsed_time(e1) / 5, 3)
print(json.dumps(out))
