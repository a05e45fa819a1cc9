"""texture_gs_b200/texture_io.py against vectors produced by the reference's own code
(tests/golden/make_texture_io_golden.py) plus layout properties."""
from pathlib import Path

import numpy as np
import pytest
import torch

from texture_gs_b200 import texture_io as TIO

Z = np.load(Path(__file__).resolve().parent / "golden" / "texture_io.npz")


def test_cube_map_and_change_texture_match_the_reference_code():
    tex, cross = torch.from_numpy(Z["texture"]), torch.from_numpy(Z["cross"])
    assert np.array_equal(TIO.cube_map(tex).numpy(), Z["cube_map"])
    for mode in (-1, 0, 1, 2, 3):
        got = TIO.change_texture(tex, cross, mode=mode).numpy()
        ref = Z[f"changed_mode{mode}"]
        fin = np.isfinite(ref)                       # mode 2 divides by the (partly zero) painted image
        assert np.array_equal(np.isfinite(got), fin), mode
        assert np.abs(got[fin] - ref[fin]).max() <= 1e-6 * max(1.0, np.abs(ref[fin]).max()), mode


def test_cross_layout_round_trip_and_resize():
    g = torch.Generator().manual_seed(1)
    rgb = torch.rand(6, 5, 5, 3, generator=g)
    tex = TIO.rgb2sh0(rgb)
    cross = TIO.cube_map(tex)
    assert cross.shape == (15, 20, 3)
    assert torch.allclose(TIO.faces_from_cross(cross), rgb, atol=1e-6)
    assert float(cross[:5, :5].abs().max()) == 0.0 and float(cross[10:, 10:].abs().max()) == 0.0       # empty corners
    assert torch.allclose(TIO.change_texture(tex, cross, mode=-1), tex, atol=1e-5)
    big = TIO.resize_cross(cross, 10)
    assert big.shape == (30, 40, 3)
    const = TIO.resize_cross(torch.full((6, 8, 3), 0.25), 7)
    assert torch.allclose(const, torch.full((21, 28, 3), 0.25))


def test_sample_cube_hits_texel_centres_and_sphere_map_directions():
    """Texel (face s, row iy, col ix) is seen in direction cube_to_dir(s, 2(ix+.5)/R-1, 2(iy+.5)/R-1)
    (NVDIFFREC/util.py:94-107); the lat-long image looks down -z at its centre column and up (+y) at its top row."""
    R = 4
    tex = torch.arange(6 * R * R * 3, dtype=torch.float64).reshape(6, R, R, 3)
    c = (2 * (torch.arange(R, dtype=torch.float64) + 0.5) / R - 1)
    y, x = torch.meshgrid(c, c, indexing="ij")
    one = torch.ones_like(x)
    dirs = [(one, -y, -x), (-one, -y, x), (x, one, y), (x, -one, -y), (x, -y, one), (-x, -y, -one)]
    for s, d in enumerate(dirs):
        got = TIO.sample_cube(tex, torch.stack(d, dim=-1) * 3.7)       # any positive scale
        assert torch.allclose(got, tex[s], atol=1e-9), s
    faces = torch.zeros(6, R, R, 3, dtype=torch.float64)
    for s in range(6):
        faces[s] = (s + 1) / 10.0
    sm = TIO.sphere_map(TIO.rgb2sh0(faces), (8, 16))
    assert sm.shape == (8, 16, 3)
    assert abs(float(sm[0, 3, 0]) - 0.3) < 1e-9          # top row: +y face (index 2)
    assert abs(float(sm[7, 3, 0]) - 0.4) < 1e-9          # bottom row: -y face (index 3)
    assert abs(float(sm[4, 8, 0]) - 0.6) < 1e-9          # centre column, gx ~ 0: direction (0, ., -1) -> -z face (index 5)
    assert abs(float(sm[4, 0, 0]) - 0.5) < 1e-9          # left edge, gx ~ -1: direction (0, ., +1) -> +z face (index 4)


def test_checkpoint_tuple_reader(tmp_path):
    """state_dict layout of models/texture_gaussian3d.py:145-171 -> the accessors uv_tex_render reads."""
    g = torch.Generator().manual_seed(2)
    N, R = 50, 4
    params = (torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g),
              torch.randn(N, 1, generator=g), torch.randn(N, 15, 3, generator=g), torch.randn(6, R, R, 3, generator=g))
    sd = dict(hyperparams=(2, 1.5), optim_state=(), net_state=({}, {}, {"weight": torch.randn(1, 128, generator=g)}), params=params)
    path = tmp_path / "30000.pth"
    torch.save((sd, 30000), path)
    gs, it = TIO.load_checkpoint(path)
    assert it == 30000 and gs.active_sh_degree == 2
    assert torch.equal(gs.get_xyz, params[0]) and torch.equal(gs.get_texture, params[5]) and torch.equal(gs.get_shs, params[4])
    assert torch.allclose(gs.get_scaling, params[1].exp())
    assert torch.allclose(gs.get_opacity, torch.sigmoid(params[3]))
    assert torch.allclose(gs.get_rotation.norm(dim=1), torch.ones(N))
    assert gs.geo_emb.shape == (128,)


def test_tcnn_flat_parameter_split():
    w_in, w_hid, w_out = torch.randn(128, 16), torch.randn(128, 128), torch.randn(16, 128)
    flat = torch.cat([w_in.reshape(-1), w_hid.reshape(-1), w_out.reshape(-1)])
    a, b, c = TIO.tcnn_mlp_weights(flat, 3, 3, n_hidden=2)
    assert torch.equal(a, w_in[:, :3]) and torch.equal(b, w_hid) and torch.equal(c, w_out[:3])
    a, b = TIO.tcnn_mlp_weights(torch.cat([w_in.reshape(-1), w_hid.reshape(-1)]), 3, 128, n_hidden=1)
    assert torch.equal(a, w_in[:, :3]) and torch.equal(b, w_hid)


def test_tcnn_checkpoint_uses_its_padded_input_columns_as_a_bias(tmp_path):
    """The shipped config trains the UV network with tiny-cuda-nn (use_tcnn: True, configs/texture_gaussian3d.yaml:20-27):
    bias-free FullyFusedMLPs whose 3 -> 128 layer has its input padded to 16 columns fed with the constant 1 — 13 columns
    that act as a learned bias. A stand-in of that network (bias-free matrices of models/modules/utils.py:29-41's shapes
    applied to the one-padded input) and the FusedUVNet loaded from the same flat parameter vectors must give the same uv
    and Jacobian; dropping the padded columns (what the loader did before) must NOT."""
    from oracle import uvnet_ref
    from texture_gs_b200.uvnet import FusedUVNet
    g = torch.Generator().manual_seed(5)
    w1p, w2 = torch.randn(128, 16, generator=g) * 0.3, torch.randn(128, 128, generator=g) * 0.1
    w3, w4, w5p = torch.randn(128, 128, generator=g) * 0.1, torch.randn(128, 128, generator=g) * 0.1, torch.randn(16, 128, generator=g) * 0.1
    emb = torch.randn(128, generator=g) * 0.1
    xyz = torch.randn(64, 3, generator=g)

    def tcnn_stand_in(x):
        xp = torch.cat([x, torch.ones(x.shape[0], 13)], dim=1)          # identity encoding: padded columns = 1
        h = torch.relu(xp @ w1p.T) @ w2.T
        h = torch.relu(h + emb[None])
        h = torch.relu(torch.relu(h @ w3.T) @ w4.T)
        return torch.nn.functional.normalize((h @ w5p.T)[:, :3], dim=-1)

    sd = dict(hyperparams=(3, 1.0), optim_state=(), params=(xyz, xyz, torch.randn(64, 4), torch.randn(64, 1), None, torch.randn(6, 2, 2, 3)),
              net_state=({"pre_mlp.params": torch.cat([w1p.reshape(-1), w2.reshape(-1)]).half(),
                          "mlp.params": torch.cat([w3.reshape(-1), w4.reshape(-1), w5p.reshape(-1)]).half()}, {}, {"weight": emb[None]}))
    net = FusedUVNet(bias=True)
    TIO.CheckpointGaussians(sd, uv_net=net)
    p = {k: v.detach() for k, v in net.state_dict().items()}
    w1h, w2h, w3h, w4h, w5h = (t.half().float() for t in (w1p, w2, w3, w4, w5p))       # the checkpoint stores halves
    w1p, w2, w3, w4, w5p = w1h, w2h, w3h, w4h, w5h
    want = tcnn_stand_in(xyz)
    got = uvnet_ref.uv_net_forward(xyz, emb, p)
    assert torch.allclose(got, want, atol=1e-5)
    jw = torch.autograd.functional.jacobian(lambda q: tcnn_stand_in(q).sum(dim=0), xyz).permute(1, 0, 2).reshape(-1, 9)
    assert torch.allclose(uvnet_ref.grad_uvs(xyz, emb, p), jw, atol=1e-4)
    assert torch.equal(p["pre_mlp.0.bias"], w1p[:, 3:].sum(dim=1)) and float(p["pre_mlp.2.bias"].abs().max()) == 0.0
    p_old = dict(p, **{"pre_mlp.0.bias": torch.zeros(128)})
    assert not torch.allclose(uvnet_ref.uv_net_forward(xyz, emb, p_old), want, atol=1e-2)
    with pytest.raises(ValueError, match="bias=True"):
        TIO.CheckpointGaussians(sd, uv_net=FusedUVNet(bias=False))
