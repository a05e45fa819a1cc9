"""Generates tests/golden/uvnet.npz by running the REFERENCE's own UVNet (models/modules/uv_net.py:8-36, built by
models/modules/utils.py:44-61 with ``use_tcnn: False`` — the nn.Linear variant; tiny-cuda-nn and addict are absent
from this image and are stubbed, they are not reached on this code path) and the Jacobian recipe of
TextureGaussian3D.get_grad_uvs (models/texture_gaussian3d.py:217-227). These vectors PIN oracle/uvnet_ref.py.
Run from the repo root inside the build container:  python tests/golden/make_uvnet_golden.py
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch
from torch.autograd.functional import jacobian

REF = Path("/root/reference/models/modules")


class _AttrDict(dict):
    __getattr__ = dict.get


def _load_reference_uvnet():
    sys.modules.setdefault("tinycudann", types.ModuleType("tinycudann"))
    addict = types.ModuleType("addict")
    addict.Dict = _AttrDict
    sys.modules.setdefault("addict", addict)
    pkg = types.ModuleType("refmodules")
    pkg.__path__ = [str(REF)]
    sys.modules["refmodules"] = pkg
    for name in ("utils", "uv_net"):
        spec = importlib.util.spec_from_file_location(f"refmodules.{name}", REF / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refmodules.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["refmodules.uv_net"].UVNet


def main():
    UVNet = _load_reference_uvnet()
    mlp = lambda layers: _AttrDict(use_tcnn=False, n_hidden_layers=layers, n_neurons=128, hash_grid_cfg=None)
    out = {}
    for tag, (n, seed, affine) in {"a": (67, 0, False), "b": (40, 1, True)}.items():
        torch.manual_seed(seed)
        cfg = _AttrDict(emb_dim=128, pre_mlp_cfg=mlp(1), mlp_cfg=mlp(2), aabb_min=None, aabb_max=None,
                        xyz_offset=[0.1, -0.2, 0.05] if affine else None, xyz_scale=[1.5, 0.8, 2.0] if affine else None)
        net = UVNet(cfg)
        xyz = torch.randn(n, 3)
        emb = 0.5 * torch.randn(128)
        uv = net(xyz, emb)
        func = lambda inputs: net(inputs, emb).float().contiguous().sum(dim=0)          # texture_gaussian3d.py:223-224
        grad_uvs = jacobian(func=func, inputs=xyz).permute(1, 0, 2).reshape(-1, 9).contiguous()
        out.update({f"{tag}_xyz": xyz.numpy(), f"{tag}_emb": emb.numpy(), f"{tag}_uv": uv.detach().numpy(), f"{tag}_grad_uvs": grad_uvs.numpy()})
        if affine:
            out[f"{tag}_offset"] = np.asarray(cfg.xyz_offset, np.float32)
            out[f"{tag}_scale"] = np.asarray(cfg.xyz_scale, np.float32)
        for k, v in net.state_dict().items():
            out[f"{tag}_p_{k}"] = v.numpy()
    dst = Path(__file__).resolve().parent / "uvnet.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, sorted(k for k in out if k.startswith("a_p_")))


if __name__ == "__main__":
    main()
