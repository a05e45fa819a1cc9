"""The scalar C oracle (oracle/raster_c.c: plain loops, hand-derived backward, OpenMP over tiles) against the torch
oracle (oracle/raster_ref.py: vectorised, autograd) — two independent restatements of spec E1-E13 that must agree to
rounding in float64, forward AND backward, including the per-pixel conditioning flags; the float32 build stays within
the parity tolerances of the float64 one. The C oracle is what makes direct (not property-based) parity checks at
BASELINE.json's full sizes affordable (tests/test_gpu_zz_fullsize.py) and is the CPU arm of bench.py."""
import shutil

import pytest
import torch

from oracle import raster_c
from util import ABS_TOL, GRAD_RTOL, oracle_settings, rel_err, run_oracle
from texture_gs_b200.scene import SyntheticGaussians, orbit_cameras, output_cotangents, sphere_shell_scene

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")


def _run_c(g, cam, bg, cot, dtype=torch.float64, scale_modifier=1.0, threads=0):
    t = g.to(dtype=dtype).tensors()
    st = oracle_settings(cam, g.active_sh_degree, dtype=dtype, bg=bg, scale_modifier=scale_modifier)
    return raster_c.rasterize(t["xyz"], t["shs"], t["opacity"], t["scaling"], t["rotation"], t["uvs"], t["grad_uvs"], t["texture"], st,
                              cotangents=cot, dtype=dtype, threads=threads)


@pytest.mark.parametrize("n,w,h,r,deg,seed,cov,sm", [(800, 64, 48, 16, 3, 0, 4.0, 0.9), (300, 70, 50, 8, 0, 4, 16.0, 1.0), (1, 17, 33, 4, 1, 1, 8.0, 1.0),
                                                     (2000, 96, 80, 64, 2, 7, 40.0, 1.3), (40, 31, 15, 1, 3, 2, 1.0, 1.0)])
def test_c_oracle_equals_torch_oracle_in_float64(n, w, h, r, deg, seed, cov, sm):
    g = sphere_shell_scene(n, r, sh_degree=deg, seed=seed, tex_seed=seed + 1, coverage=cov)
    cam = orbit_cameras(1, w, h, seed=seed + 2)[0]
    bg = (0.2, 0.4, 0.6)
    cot = output_cotangents(h, w, seed=3)
    ref, aux, gref = run_oracle(g, cam, bg=bg, cot=cot, dtype=torch.float64, scale_modifier=sm)
    img, dep, nrm, alp, radii, caux = _run_c(g, cam, bg, cot, scale_modifier=sm)
    for name, a, b in zip(("image", "depth", "norm", "alpha"), (img, dep, nrm, alp), ref[:4]):
        assert float((a - b).abs().max()) <= 1e-11, name
    assert torch.equal(radii, ref[4]) and torch.equal(caux["n_contrib"].long(), aux["n_contrib"].long())
    assert float((caux["final_T"] - aux["final_T"]).abs().max()) <= 1e-12
    assert (caux["num_pairs"], caux["num_visible"], caux["num_blend"]) == (aux["num_pairs"], aux["num_visible"], aux["num_blend"])
    for k, rg in gref.items():
        if rg is None:
            continue
        assert rel_err(caux["grads"][k].reshape(rg.shape), rg) <= 1e-9, k
    # the conditioning flags are the same sets up to the pixels whose quantity sits within rounding of the flag's own threshold
    for key in ("ambiguous", "grad_ambiguous", "grazing", "texel_boundary"):
        assert float((caux[key] != aux[key]).float().mean()) <= 2e-3, key


def test_c_oracle_thread_count_does_not_change_the_outputs_and_float32_stays_within_tolerance():
    g = sphere_shell_scene(1500, 32, sh_degree=3, seed=11, tex_seed=12)
    cam = orbit_cameras(1, 112, 80, seed=13)[0]
    bg = (0.1, 0.2, 0.3)
    o1 = _run_c(g, cam, bg, None, threads=1)
    o4 = _run_c(g, cam, bg, None, threads=4)
    for a, b in zip(o1[:5], o4[:5]):
        assert torch.equal(a, b)                              # the forward has no cross-thread accumulation
    keep = (~o1[5]["grad_ambiguous"]).double()
    cot = [c.double() * keep for c in output_cotangents(80, 112, seed=14)]
    g64 = _run_c(g, cam, bg, cot)[5]["grads"]
    g64b = _run_c(g, cam, bg, cot, threads=3)[5]["grads"]
    for k, v in g64.items():
        if v is not None:
            assert rel_err(g64b[k], v) <= 1e-12, k            # atomics reorder float64 sums only
    f32 = _run_c(g, cam, bg, cot, dtype=torch.float32)
    amb = o1[5]["ambiguous"]
    for i, name in enumerate(("image", "depth", "norm", "alpha")):
        d = (f32[i].double() - o1[i]).abs().amax(0)
        assert float(d[~amb].max()) <= ABS_TOL * (3.0 if name == "depth" else 1.0), name
    for k, v in g64.items():
        if v is not None:
            assert rel_err(f32[5]["grads"][k].double(), v) <= GRAD_RTOL, k


def test_c_oracle_edge_cases_empty_culled_and_degenerate_inputs():
    cam = orbit_cameras(1, 33, 17, seed=3)[0]
    bg = (0.3, 0.5, 0.7)
    g0 = sphere_shell_scene(4, 4, sh_degree=0)
    t = {k: (v[:0] if (v is not None and k != "texture") else v) for k, v in g0.tensors().items()}
    ge = SyntheticGaussians(active_sh_degree=0, **{k: (v.detach() if v is not None else None) for k, v in t.items()})
    img, dep, nrm, alp, radii, aux = _run_c(ge, cam, bg, output_cotangents(17, 33, seed=1))
    assert torch.allclose(img, torch.tensor(bg, dtype=torch.float64)[:, None, None].expand(3, 17, 33)) and float(alp.abs().max()) == 0.0
    assert radii.numel() == 0 and aux["num_pairs"] == 0 and float(aux["grads"]["texture"].abs().max()) == 0.0
    gb = sphere_shell_scene(64, 4, sh_degree=0)
    tb = {k: (v.detach().clone() if v is not None else None) for k, v in gb.tensors().items()}
    tb["xyz"][:32] = cam.camera_center * 2.0                    # behind the camera: culled
    tb["xyz"][40] = float("nan")
    tb["rotation"][41] = 0.0
    tb["scaling"][42] = 0.0
    gg = SyntheticGaussians(active_sh_degree=0, **tb)
    ref, raux, gref = run_oracle(gg, cam, bg=bg, dtype=torch.float64)
    img, dep, nrm, alp, radii, aux = _run_c(gg, cam, bg, output_cotangents(17, 33, seed=1))
    assert int(radii[:32].max()) == 0 and int(radii[40]) == 0 and torch.equal(radii, ref[4])
    assert float((img - ref[0]).abs().max()) <= 1e-11 and float((alp - ref[3]).abs().max()) <= 1e-11
    assert all(bool(torch.isfinite(v).all()) for v in aux["grads"].values() if v is not None)
    assert float(aux["grads"]["xyz"][:32].abs().max()) == 0.0
