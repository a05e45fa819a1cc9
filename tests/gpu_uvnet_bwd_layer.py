"""tcgen05 backward-layer kernel (texgs_uvmlp_backward_layer) against torch on the GPU: python tests/gpu_uvnet_bwd_layer.py [N ...]
gW = delta^T a, delta_out = (delta W) * (a > 0), colsum = sum_rows(delta_out); fp16 operands, fp32 accumulation."""
import ctypes as C, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from texture_gs_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda")
out = []
for N in [int(x) for x in sys.argv[1:]] or [128, 1000, 4096, 500_000]:
    g = torch.Generator(device="cpu").manual_seed(N)
    delta = (torch.randn(N, 128, generator=g) * 0.5).half().to(dev)
    a = torch.relu(torch.randn(N, 128, generator=g)).half().to(dev)
    W = (torch.randn(128, 128, generator=g) * 0.1).half().to(dev)           # [out][in]
    Wt = W.T.contiguous()
    d_out = torch.full((N, 128), float("nan"), dtype=torch.float16, device=dev)
    gW = torch.zeros(128, 128, device=dev)
    cs = torch.zeros(128, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.texgs_uvmlp_backward_layer(N, p(delta), p(a), p(Wt), p(d_out), p(gW), p(cs), st), "texgs_uvmlp_backward_layer")
    torch.cuda.synchronize()
    ref_gW = delta.float().T @ a.float()
    ref_pre = delta.float() @ W.float()
    ref_d = (ref_pre * (a > 0)).half()
    ref_cs = ref_d.float().sum(0)
    e_gW = float((gW - ref_gW).abs().max() / ref_gW.abs().max())
    e_d = float((d_out.float() - ref_d.float()).abs().max() / ref_d.float().abs().max())
    e_cs = float((cs - ref_cs).abs().max() / ref_cs.abs().max())
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = None
    if N >= 100_000:
        for _ in range(3):
            lib.texgs_uvmlp_backward_layer(N, p(delta), p(a), p(Wt), p(d_out), p(gW), p(cs), st)
        t0.record()
        for _ in range(10):
            lib.texgs_uvmlp_backward_layer(N, p(delta), p(a), p(Wt), p(d_out), p(gW), p(cs), st)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        t0.record()
        for _ in range(10):
            x = torch.mm(delta.T, a, out_dtype=torch.float32); y = torch.mm(delta, W)
        t1.record(); torch.cuda.synchronize()
        ms = (ms, t0.elapsed_time(t1) / 10)
    out.append({"N": N, "err_gW": e_gW, "err_delta": e_d, "err_colsum": e_cs, "ms (kernel, 2 torch.mm)": ms, "nan": bool(torch.isnan(d_out).any())})
print(json.dumps(out))
ok = all(o["err_gW"] < 2e-3 and o["err_delta"] < 2e-3 and o["err_colsum"] < 2e-3 and not o["nan"] for o in out)
sys.exit(0 if ok else 1)
