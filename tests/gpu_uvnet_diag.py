"""Bring-up diagnostics of the tcgen05 UV-MLP kernel: raw layer accumulators of the first tile against torch matmuls of
the same fp16 operands, for both shared-memory descriptor conventions. Prints one JSON line."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import uvnet_ref as UR
from texture_gs_b200.uvnet import FusedUVNet, _UvMlp

torch.manual_seed(0)
N = 100
p = UR.random_params(seed=3, bias=True)
net = FusedUVNet(bias=True).cuda()
net.load_state_dict(p)
xyz = torch.randn(N, 3).cuda()
emb = (0.5 * torch.randn(128)).cuda()
out = {}
for mode in (True,):
    try:
        uv, jac = net._call(xyz, emb, debug=mode)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        out[str(mode)] = {"error": str(e)[:200]}
        continue
    dbg = _UvMlp.last_debug.cpu()
    acc1 = dbg[:128 * 128].view(128, 128)                 # layer-1 accumulators (hi/lo-split K=16 MMA): rows = stream*32 + point
    acc2 = dbg[128 * 128:2 * 128 * 128].view(128, 128)    # layer-2 accumulators
    # expected operands: a1 rows (value: relu(W1 x + b1); tangent j: W1[:, j] * (pre > 0)), rounded to fp16
    x = xyz[:32].cpu()
    pre = x @ p["pre_mlp.0.weight"].T + p["pre_mlp.0.bias"]
    rows = [torch.relu(pre)] + [(pre > 0).float() * p["pre_mlp.0.weight"][:, j][None, :] for j in range(3)]
    exp1 = torch.cat([pre] + [p["pre_mlp.0.weight"][:, j][None, :].expand(32, 128) for j in range(3)], 0)
    A = torch.cat(rows, 0).half().float()
    W2 = p["pre_mlp.2.weight"].half().float()
    exp2 = A @ W2.T
    uv_ref = UR.uv_net_forward(xyz.cpu().double(), emb.cpu().double(), {k: v.double() for k, v in p.items()})
    j_ref = UR.grad_uvs(xyz.cpu().double(), emb.cpu().double(), {k: v.double() for k, v in p.items()})
    out[str(mode)] = {"layer1_acc_max_err": float((acc1 - exp1).abs().max()), "layer2_acc_max_err": float((acc2 - exp2).abs().max()), "layer2_acc_max": float(exp2.abs().max()),
                      "acc_row0": [round(float(v), 4) for v in acc2[0, :4]], "exp_row0": [round(float(v), 4) for v in exp2[0, :4]],
                      "uv_max_err": float((uv.cpu().double() - uv_ref).abs().max()),
                      "jac_max_err": float((jac.cpu().double() - j_ref).abs().max()), "jac_max": float(j_ref.abs().max())}
print(json.dumps(out))
