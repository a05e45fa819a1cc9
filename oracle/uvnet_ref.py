"""ORACLE (test infrastructure, NOT product code) — restatement of the reference's UV + Jacobian producer.

  uv_net_forward — reference models/modules/uv_net.py:19-36 (UVNet.forward) over the nn.Linear networks that
                   models/modules/utils.py:44-55 (build_nn_network) builds for ``use_tcnn: False``:
                   pre_mlp = Linear(3,128) ReLU Linear(128,emb);  x = relu(pre_mlp(x') + emb);
                   mlp = Linear(emb,128) ReLU Linear(128,128) ReLU Linear(128,3);  F.normalize(., dim=-1)
  grad_uvs       — reference models/texture_gaussian3d.py:217-227 (get_grad_uvs):
                   jacobian(lambda x: uv_net(x, emb).sum(0), xyz) -> (3,N,3) -> permute(1,0,2).reshape(-1,9)

PINNED for the nn.Linear variant: tests/golden/uvnet.npz is produced by the reference's own UVNet class
(tests/golden/make_uvnet_golden.py imports models/modules/uv_net.py with ``use_tcnn: False``). The shipped config
uses tiny-cuda-nn's FullyFusedMLP (fp16, no biases; configs/texture_gaussian3d.yaml:18-27), which is absent from
this image: for that variant the oracle is this same function with ``bias=None`` — same algebra, unpinned numerics.

  forward_mode_fp16 — the same network evaluated the way the CUDA kernel evaluates it: value and three forward-mode
                   tangents, hidden/output weights and every activation rounded to fp16, fp32 accumulation. The
                   Jacobian of a ReLU network is piecewise constant in the masks, so a unit whose pre-activation is
                   within fp16 rounding of 0 changes J by O(1/width) between ANY two precisions (the reference's own
                   fp16 tiny-cuda-nn path included); parity of J and of the gradients is therefore asserted tightly
                   against this matched-precision form and statistically against the fp64 form.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch.autograd.functional import jacobian

KEYS = ("pre_mlp.0", "pre_mlp.2", "mlp.0", "mlp.2", "mlp.4")       # nn.Sequential indices of the Linear layers


def random_params(seed: int = 0, bias: bool = True, emb_dim: int = 128, hidden: int = 128, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """nn.Linear-style initialisation (uniform +-1/sqrt(fan_in)), deterministic."""
    g = torch.Generator().manual_seed(seed)
    dims = {"pre_mlp.0": (hidden, 3), "pre_mlp.2": (emb_dim, hidden), "mlp.0": (hidden, emb_dim), "mlp.2": (hidden, hidden), "mlp.4": (3, hidden)}
    out = {}
    for k, (o, i) in dims.items():
        bound = 1.0 / (i ** 0.5)
        out[k + ".weight"] = ((torch.rand(o, i, generator=g) * 2 - 1) * bound).to(dtype)
        if bias:
            out[k + ".bias"] = ((torch.rand(o, generator=g) * 2 - 1) * bound).to(dtype)
    return out


def _lin(x, p, k):
    y = x @ p[k + ".weight"].T
    b = p.get(k + ".bias")
    return y if b is None else y + b


def uv_net_forward(xyz: torch.Tensor, emb: torch.Tensor, p: Dict[str, torch.Tensor], xyz_offset: Optional[torch.Tensor] = None,
                   xyz_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    x = xyz
    if xyz_offset is not None and xyz_scale is not None:
        x = (x - xyz_offset) / xyz_scale
    h = _lin(torch.relu(_lin(x, p, "pre_mlp.0")), p, "pre_mlp.2")
    h = torch.relu(h + emb[None, :])
    h = torch.relu(_lin(h, p, "mlp.0"))
    h = torch.relu(_lin(h, p, "mlp.2"))
    return F.normalize(_lin(h, p, "mlp.4"), dim=-1)


def grad_uvs(xyz: torch.Tensor, emb: torch.Tensor, p, xyz_offset=None, xyz_scale=None) -> torch.Tensor:
    j = jacobian(lambda inp: uv_net_forward(inp, emb, p, xyz_offset, xyz_scale).sum(dim=0), xyz)      # (3, N, 3)
    return j.permute(1, 0, 2).reshape(-1, 9).contiguous()


def _ste_half(t: torch.Tensor) -> torch.Tensor:
    """Round to fp16 (straight-through for autograd)."""
    return t + (t.detach().half().to(t.dtype) - t.detach())


def forward_mode_fp16(xyz: torch.Tensor, emb: torch.Tensor, p: Dict[str, torch.Tensor], xyz_offset=None, xyz_scale=None):
    """(uv (N,3), grad_uvs (N,9)) with the kernel's rounding points; fp32 tensors in, differentiable (uv only)."""
    f = torch.float32
    inv_scale = torch.ones(3, dtype=f) if xyz_scale is None else (1.0 / xyz_scale.to(f))
    x = xyz.to(f) if xyz_offset is None else (xyz.to(f) - xyz_offset.to(f)) * inv_scale
    W1, b1 = p["pre_mlp.0.weight"].to(f), p.get("pre_mlp.0.bias")
    pre = x @ W1.T + (0 if b1 is None else b1.to(f))
    m = (pre > 0).to(f)
    a = _ste_half(torch.relu(pre))
    T = [(m * (W1[:, j] * inv_scale[j])[None, :]).half().to(f) for j in range(3)]
    for k, add_emb in (("pre_mlp.2", True), ("mlp.0", False), ("mlp.2", False)):
        W = _ste_half(p[k + ".weight"].to(f))
        b = p.get(k + ".bias")
        pre = a @ W.T + (0 if b is None else b.to(f)) + (emb.to(f)[None, :] if add_emb else 0)
        m = (pre > 0).to(f)
        a = _ste_half(torch.relu(pre))
        T = [((t @ W.detach().T) * m).half().to(f) for t in T]
    W5 = _ste_half(p["mlp.4.weight"].to(f))
    b5 = p.get("mlp.4.bias")
    out = a @ W5.T + (0 if b5 is None else b5.to(f))
    inv_len = 1.0 / out.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    uv = out * inv_len
    cols = []
    for t in T:
        t5 = t @ W5.detach().T
        cols.append((t5 - uv.detach() * (uv.detach() * t5).sum(-1, keepdim=True)) * inv_len.detach())
    J = torch.stack(cols, dim=-1)                                      # (N, 3 [i], 3 [j])
    return uv, J.reshape(-1, 9).contiguous()
