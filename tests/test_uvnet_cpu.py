"""CPU checks of the N1 oracle (oracle/uvnet_ref.py) against the reference-generated golden vectors, and host logic."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import uvnet_ref as UR


def test_uvnet_oracle_matches_the_reference_code_golden_vectors():
    z = np.load(Path(__file__).resolve().parent / "golden" / "uvnet.npz")
    for tag in ("a", "b"):
        p = {k[len(tag) + 3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}_p_")}
        off = torch.from_numpy(z[f"{tag}_offset"]) if f"{tag}_offset" in z.files else None
        sc = torch.from_numpy(z[f"{tag}_scale"]) if f"{tag}_scale" in z.files else None
        xyz, emb = torch.from_numpy(z[f"{tag}_xyz"]), torch.from_numpy(z[f"{tag}_emb"])
        assert float((UR.uv_net_forward(xyz, emb, p, off, sc) - torch.from_numpy(z[f"{tag}_uv"])).abs().max()) < 1e-6
        assert float((UR.grad_uvs(xyz, emb, p, off, sc) - torch.from_numpy(z[f"{tag}_grad_uvs"])).abs().max()) < 1e-5


def test_oracle_jacobian_is_tangent_to_the_sphere_and_matches_finite_differences():
    p = {k: v.double() for k, v in UR.random_params(seed=1, bias=True).items()}
    g = torch.Generator().manual_seed(0)
    xyz, emb = torch.randn(20, 3, generator=g).double(), torch.randn(128, generator=g).double()
    uv, J = UR.uv_net_forward(xyz, emb, p), UR.grad_uvs(xyz, emb, p).view(-1, 3, 3)
    assert float((J * uv[:, :, None]).sum(1).abs().max()) < 1e-12
    h = 1e-6
    for j in range(3):
        d = torch.zeros(3, dtype=torch.float64); d[j] = h
        fd = (UR.uv_net_forward(xyz + d, emb, p) - UR.uv_net_forward(xyz - d, emb, p)) / (2 * h)
        assert float((fd - J[:, :, j]).abs().max()) < 1e-6


def test_fused_module_mirrors_reference_state_dict_and_refuses_cpu():
    from texture_gs_b200.uvnet import FusedUVNet
    net = FusedUVNet(bias=True)
    assert set(net.state_dict().keys()) == set(UR.random_params(bias=True).keys())
    assert set(FusedUVNet(bias=False).state_dict().keys()) == set(UR.random_params(bias=False).keys())
    with pytest.raises(RuntimeError):
        net(torch.zeros(4, 3), torch.zeros(128))
    with pytest.raises(ValueError):
        FusedUVNet(xyz_offset=[0, 0, 0])
