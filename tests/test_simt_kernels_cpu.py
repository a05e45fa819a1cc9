"""CPU-only parity of the REAL kernel source: the CUDA files of texture_gs_b200/csrc are compiled with g++ on top of
the SIMT emulator in tests/simt (every CUDA thread a fiber; warp collectives, block barriers, mbarrier / bulk-copy
pipelines and atomics emulated, see simt_emu.h) and driven through the same C-ABI call sequence as the product.

These tests do not replace the ``-m gpu`` parity tests (the emulator says nothing about PTX semantics, memory ordering
between real warps or performance); they check on a box without a GPU that the kernels' arithmetic, list handling and
warp choreography reproduce the oracle — with the tolerances of the GPU suite (BASELINE.json: 1e-4 abs, 1e-3 rel).
Sizes are moderate: the emulator runs about 1e6 warp collectives per second (a full 500 k / 1080p forward + backward
takes three minutes: tools/emu_check.py).
"""
import math
import shutil

import numpy as np
import pytest
import torch

from util import ABS_TOL, GRAD_RTOL, check_backward, check_forward, compare_images, oracle_settings, rel_err, run_emu, run_oracle
from texture_gs_b200.scene import SyntheticGaussians, orbit_cameras, output_cotangents, sphere_shell_scene

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the SIMT emulator needs g++")


@pytest.fixture(scope="module")
def emu():
    from simt import emu as E
    E.build()
    return E


def _cam_kw(cam, bg, sh_degree):
    return dict(H=cam.image_height, W=cam.image_width, tanfovx=math.tan(cam.FoVx / 2), tanfovy=math.tan(cam.FoVy / 2), bg=bg,
                viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center, sh_degree=sh_degree)


@pytest.mark.parametrize("n,w,h,r,deg,seed,eager,noise", [(1500, 96, 64, 32, 3, 0, False, 0), (1500, 96, 64, 32, 3, 0, True, 16),
                                                         (400, 70, 50, 16, 0, 4, False, 0)])
def test_emulated_kernels_match_oracle_forward_and_backward(emu, n, w, h, r, deg, seed, eager, noise):
    """The whole pipeline (preprocess, scan, scatter, per-tile sort, render forward / backward, preprocess backward)
    against the oracle; the last case has an image size that is not a multiple of the tile and no SH. ``eager``: the
    records' bulk copies land at issue instead of as late as legal — the two extreme timings of the 2-stage ring.
    ``noise``: every fast-math intrinsic (__expf, __fdividef, __logf, rsqrtf) returns its result with a pseudo-random
    relative error of up to 16 ulp (2e-6), more than the GPU's approximate units have: the 1e-4 / 1e-3 tolerances and
    the oracle's threshold-proximity flags must absorb that."""
    g = sphere_shell_scene(n, r, sh_degree=deg, seed=seed, tex_seed=seed + 1)
    cam = orbit_cameras(1, w, h, seed=seed + 2)[0]
    emu.build().simt_set_eager_copies(1 if eager else 0)
    emu.build().simt_set_fastmath_noise(noise)
    try:
        check_forward(g, cam, bg=(0.2, 0.4, 0.6), runner=run_emu, max_amb=0.3)
        check_backward(g, cam, bg=(0.2, 0.4, 0.6), runner=run_emu, uv_tol=GRAD_RTOL if r > 16 else 5e-3, max_flag=0.3)
    finally:
        emu.build().simt_set_eager_copies(0)
        emu.build().simt_set_fastmath_noise(0)


def test_emulated_binning_is_an_order_preserving_subsequence_of_the_oracle_lists(emu):
    """Sorted per-tile lists == the oracle's (tile, depth, index) order (spec E4) minus pairs that cannot blend."""
    g = sphere_shell_scene(1500, 16, sh_degree=0, seed=2)
    cam = orbit_cameras(1, 96, 80, seed=5)[0]
    _, stats, _ = run_emu(g, cam)
    res = run_emu.last
    _, aux, _ = run_oracle(g, cam)
    offs = res.tile_offset.numpy().astype(np.int64)
    ids = res.sorted_ids.numpy().astype(np.int64)
    K = stats.num_pairs
    assert offs[-1] == K and K <= aux["num_pairs"]
    tile_of = np.repeat(np.arange(offs.shape[0] - 1), np.diff(offs))
    ref_key = aux["tile_of"].astype(np.int64) * (1 << 32) + aux["gid_of"].astype(np.int64)
    got_key = tile_of * (1 << 32) + ids
    pos = {k: i for i, k in enumerate(ref_key.tolist())}
    idx = np.array([pos[k] for k in got_key.tolist()])          # KeyError = a pair the spec does not have
    assert (np.diff(idx) > 0).all()
    kept = np.zeros(ref_key.shape[0], dtype=bool)
    kept[idx] = True
    assert kept[aux["pair_contributes"]].all()
    assert K < aux["num_pairs"]


def test_emulated_plain_texture_layout_dual_render_and_accumulation(emu):
    """(a) the (6,R,R,3) kernels == the packed (6,R,R,4) kernels; (b) the dual render's second image and its cotangent
    == the oracle's two renders; (c) accumulate_mask / zero_texture_grad = 0 add onto what the buffers
    hold (the fused-bucket path of GradBucket) instead of overwriting."""
    from oracle import raster_ref as RR
    N, W, H, R = 1200, 80, 64, 32
    g = sphere_shell_scene(N, R, sh_degree=3, seed=21, tex_seed=22)
    cam = orbit_cameras(1, W, H, seed=23)[0]
    bgc = (0.2, 0.1, 0.3)
    _, aux0, _ = run_oracle(g, cam, bg=bgc)
    keep = (~aux0["grad_ambiguous"]).float()
    cot = [c * keep for c in output_cotangents(H, W, seed=24)]
    cot2 = torch.randn(3, H, W, generator=torch.Generator().manual_seed(5)) * keep
    t = g.to(dtype=torch.float32, requires_grad=True).tensors()
    kw = dict(means3D=t["xyz"], opacities=t["opacity"], scales=t["scaling"], rotations=t["rotation"], shs=t["shs"], uvs=t["uvs"],
              gradient_uvs=t["grad_uvs"], texture=t["texture"], **_cam_kw(cam, bgc, 3))
    a = emu.rasterize(cotangents=cot, **kw)
    b = emu.rasterize(cotangents=cot, packed_texture=False, packed_texture_grad=False, **kw)
    for x, y in zip((a.image, a.depth, a.norm, a.alpha), (b.image, b.depth, b.norm, b.alpha)):
        assert torch.equal(x, y)
    for k in a.grads:
        assert rel_err(a.grads[k], b.grads[k]) < 1e-5, k
    # (b) dual
    st = oracle_settings(cam, 3, bg=bgc)
    o = RR.rasterize(t["xyz"], None, t["shs"], t["opacity"], t["scaling"], t["rotation"], t["uvs"], t["grad_uvs"], t["texture"], st,
                     return_aux=True, dual_no_sh=True)
    (sum((x * y).sum() for x, y in zip(o[:4], cot)) + (o[-1]["image_no_sh"] * cot2).sum()).backward()
    d = emu.rasterize(cotangents=cot, dual_no_sh=True, cot_nosh=cot2, **kw)
    amb = aux0["ambiguous"]
    assert float((d.image - o[0].detach()).abs().amax(0)[~amb].max()) <= ABS_TOL
    assert float((d.image_nosh - o[-1]["image_no_sh"].detach()).abs().amax(0)[~amb].max()) <= ABS_TOL
    names = {"means3D": "xyz", "opacities": "opacity", "scales": "scaling", "rotations": "rotation", "shs": "shs", "uvs": "uvs", "texture": "texture"}
    for k, ok in names.items():
        e = rel_err(d.grads[k].reshape(t[ok].grad.shape), t[ok].grad)
        assert e <= GRAD_RTOL, (k, e)
    # (c) accumulation on top of a constant
    c = emu.rasterize(cotangents=cot, accumulate_onto=0.25, **kw)
    for k in names:
        assert rel_err(c.grads[k] - 0.25, a.grads[k]) < 1e-4, k
    assert rel_err(c.grads["means2D"], a.grads["means2D"]) < 1e-6          # never accumulated: overwritten


def test_emulated_plain_3dgs_modes_and_cov3d_match_oracle(emu):
    """texture=None (the diff_gauss surface, render/render.py:75-84): colours from colors_precomp or full SH, and
    ``cov3Ds_precomp`` in place of scales + rotations."""
    from oracle import raster_ref as RR
    N, W, H = 900, 80, 64
    g = sphere_shell_scene(N, 4, sh_degree=3, seed=8)
    cam = orbit_cameras(1, W, H, seed=9)[0]
    gen = torch.Generator().manual_seed(1)
    cols = torch.rand(N, 3, generator=gen)
    shs_full = torch.cat([torch.randn(N, 1, 3, generator=gen), 0.1 * torch.randn(N, 15, 3, generator=gen)], dim=1)
    bgc = (0.1, 0.1, 0.3)
    st = oracle_settings(cam, 3, bg=bgc)
    for kind in ("precomp", "sh", "cov"):
        t = g.to(dtype=torch.float32, requires_grad=True).tensors()
        c_ref = cols.clone().requires_grad_(True)
        s_ref = shs_full.clone().requires_grad_(True)
        cov = None
        if kind == "cov":
            Lm = RR.quat_to_rot(t["rotation"].detach().double()) * t["scaling"].detach().double()[:, None, :]
            S = Lm @ Lm.transpose(1, 2)
            cov = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=-1).float().requires_grad_(True)
        with torch.no_grad():
            aux0 = RR.rasterize(t["xyz"], None, s_ref if kind == "sh" else None, t["opacity"], None if cov is not None else t["scaling"],
                                None if cov is not None else t["rotation"], None, None, None, st,
                                colors_precomp=None if kind == "sh" else c_ref, cov3Ds_precomp=cov, return_aux=True)[-1]
        keep = (~aux0["grad_ambiguous"]).float()
        cot = [c * keep for c in output_cotangents(H, W, seed=2)]
        if kind == "cov":
            cot[2] = torch.zeros_like(cot[2])        # the normal carries no gradient in this mode
        o = RR.rasterize(t["xyz"], None, s_ref if kind == "sh" else None, t["opacity"], None if cov is not None else t["scaling"],
                         None if cov is not None else t["rotation"], None, None, None, st,
                         colors_precomp=None if kind == "sh" else c_ref, cov3Ds_precomp=cov, return_aux=True)
        sum((x * y).sum() for x, y in zip(o[:4], cot)).backward()
        r = emu.rasterize(means3D=t["xyz"], opacities=t["opacity"], scales=None if cov is not None else t["scaling"],
                          rotations=None if cov is not None else t["rotation"], shs=s_ref if kind == "sh" else None,
                          colors_precomp=None if kind == "sh" else c_ref, cov3Ds_precomp=cov, cotangents=cot, **_cam_kw(cam, bgc, 3))
        rep = compare_images((r.image, r.depth, r.norm, r.alpha), [x.detach() for x in o[:4]], aux0["ambiguous"])
        for nme in ("image", "depth", "norm", "alpha"):
            assert rep[nme]["max_clear"] <= ABS_TOL * (3.0 if nme == "depth" else 1.0), (kind, nme, rep[nme])
        assert int((r.radii != o[4]).sum()) == 0
        pairs = [("means3D", t["xyz"].grad), ("opacities", t["opacity"].grad)]
        pairs += [("cov3Ds_precomp", cov.grad)] if kind == "cov" else [("scales", t["scaling"].grad), ("rotations", t["rotation"].grad)]
        pairs += [("shs", s_ref.grad)] if kind == "sh" else [("colors_precomp", c_ref.grad)]
        for name, ref in pairs:
            e = rel_err(r.grads[name].reshape(ref.shape), ref)
            assert e <= (3e-3 if name == "cov3Ds_precomp" else GRAD_RTOL), (kind, name, e)


def test_emulated_extra_attrs_match_oracle(emu):
    """``extra_attrs`` (P,E) -> ``extra`` (E,H,W) with E = 11 (two channel groups of the kernel, ragged tail), forward
    and backward including the alpha-chain share of the extra cotangent."""
    from oracle import raster_ref as RR
    N, W, H, R, E = 600, 48, 48, 16, 11
    g = sphere_shell_scene(N, R, sh_degree=2, seed=51, tex_seed=52)
    cam = orbit_cameras(1, W, H, seed=53)[0]
    gen = torch.Generator().manual_seed(54)
    ex0 = torch.randn(N, E, generator=gen)
    bgc = (0.2, 0.1, 0.3)
    _, aux0, _ = run_oracle(g, cam, bg=bgc)
    keep = (~aux0["grad_ambiguous"]).float()
    cot = [c * keep for c in output_cotangents(H, W, seed=55)]
    cot_e = torch.randn(E, H, W, generator=gen) * keep
    t = g.to(dtype=torch.float32, requires_grad=True).tensors()
    ex = ex0.clone().requires_grad_(True)
    o = RR.rasterize(t["xyz"], None, t["shs"], t["opacity"], t["scaling"], t["rotation"], t["uvs"], t["grad_uvs"], t["texture"],
                     oracle_settings(cam, 2, bg=bgc), extra_attrs=ex)
    (sum((x * y).sum() for x, y in zip(o[:4], cot)) + (o[5] * cot_e).sum()).backward()
    r = emu.rasterize(means3D=t["xyz"], opacities=t["opacity"], scales=t["scaling"], rotations=t["rotation"], shs=t["shs"], uvs=t["uvs"],
                      gradient_uvs=t["grad_uvs"], texture=t["texture"], extra_attrs=ex, cotangents=cot, cot_extra=cot_e, **_cam_kw(cam, bgc, 2))
    rep = compare_images((r.image, r.depth, r.norm, r.alpha, r.extra), [x.detach() for x in o[:4]] + [o[5].detach()], aux0["ambiguous"],
                         names=("image", "depth", "norm", "alpha", "extra"))
    for nme in ("image", "depth", "norm", "alpha", "extra"):
        assert rep[nme]["max_clear"] <= ABS_TOL * (3.0 if nme in ("depth", "extra") else 1.0), (nme, rep[nme])
    for name, ref in (("extra_attrs", ex.grad), ("opacities", t["opacity"].grad), ("means3D", t["xyz"].grad), ("scales", t["scaling"].grad),
                      ("rotations", t["rotation"].grad), ("uvs", t["uvs"].grad), ("texture", t["texture"].grad)):
        e = rel_err(r.grads[name].reshape(ref.shape), ref)
        assert e <= (5e-3 if name == "uvs" else GRAD_RTOL), (name, e)


def test_emulated_edge_cases_empty_culled_single_and_overflow_retry(emu):
    bgc = (0.3, 0.5, 0.7)
    cam = orbit_cameras(1, 33, 17, seed=3)[0]
    # (a) zero Gaussians: background only, nothing launched per Gaussian
    g0 = sphere_shell_scene(4, 4, sh_degree=0)
    t = {k: (v[:0] if (v is not None and k != "texture") else v) for k, v in g0.tensors().items()}
    ge = SyntheticGaussians(active_sh_degree=0, **{k: (v.detach() if v is not None else None) for k, v in t.items()})
    out, stats, grads = run_emu(ge, cam, bg=bgc, cot=output_cotangents(17, 33, seed=1))
    assert torch.allclose(out[0], torch.tensor(bgc)[:, None, None].expand(3, 17, 33))
    assert float(out[3].abs().max()) == 0.0 and stats.num_pairs == 0 and float(grads["texture"].abs().max()) == 0.0
    # (b) everything behind the camera: culled, radii == 0, zero gradients
    gb = sphere_shell_scene(64, 4, sh_degree=0)
    tb = gb.tensors()
    far = SyntheticGaussians(active_sh_degree=0, **{**{k: (v.detach() if v is not None else None) for k, v in tb.items()},
                                                      "xyz": tb["xyz"].detach() * 0 + cam.camera_center * 2.0})
    out, stats, grads = run_emu(far, cam, bg=bgc, cot=output_cotangents(17, 33, seed=1))
    assert int(out[4].max()) == 0 and float(out[3].abs().max()) == 0.0 and stats.num_visible == 0
    assert float(grads["xyz"].abs().max()) == 0.0
    # (c) a single Gaussian on a ragged image, against the oracle
    g1 = sphere_shell_scene(1, 8, sh_degree=1, seed=1, coverage=8.0)
    c1 = orbit_cameras(1, 17, 33, seed=1)[0]
    check_forward(g1, c1, bg=bgc, max_amb=0.5, runner=run_emu)
    # (d) pair capacity too small on the first attempt: the kernels self-disable, the retry is transparent
    g = sphere_shell_scene(600, 16, sh_degree=0, seed=1)
    cam2 = orbit_cameras(1, 64, 64, seed=2)[0]
    o1, s1, _ = run_emu(g, cam2)
    o2, s2, _ = run_emu(g, cam2, pair_capacity=64)
    assert s2.num_pairs == s1.num_pairs and s2.pair_capacity >= s2.num_pairs > 64
    for x, y in zip(o1, o2):
        assert torch.equal(x, y)


def test_emulated_long_lists_take_the_large_sort_kernel_and_early_termination(emu):
    """One 16x16 tile that sees ~700 splats (> 512: the 256-thread sort kernel; 22 chunks through the 2-stage ring) with
    opacities high enough that pixels stop early (n_contrib < list length: the backward starts mid-list)."""
    n = 700
    g = sphere_shell_scene(n, 8, sh_degree=0, seed=5, coverage=4.0)
    t = g.tensors()
    cam = orbit_cameras(1, 16, 16, seed=6)[0]
    c = cam.camera_center / cam.camera_center.norm()
    gen = torch.Generator().manual_seed(0)
    xyz = c[None, :] * 1.0 + 0.02 * torch.randn(n, 3, generator=gen)
    uv = xyz / xyz.norm(dim=1, keepdim=True)
    gg = SyntheticGaussians(active_sh_degree=0, **{**{k: (v.detach() if v is not None else None) for k, v in t.items()},
                                                    "xyz": xyz, "uvs": uv, "opacity": torch.full((n, 1), 0.5)})
    ref, aux, _ = run_oracle(gg, cam)
    got, stats, _ = run_emu(gg, cam)
    assert stats.max_tile_len > 512
    stopped = aux["final_T"] < 2e-4
    assert bool(stopped.any()) and int(aux["n_contrib"][stopped].min()) < stats.max_tile_len     # the stop rule really fires
    rep = compare_images(got[:4], ref[:4], aux["ambiguous"])
    for nme in ("image", "alpha"):
        assert rep[nme]["max_clear"] <= 2 * ABS_TOL, rep
    check_backward(gg, cam, runner=run_emu, max_flag=0.9, uv_tol=5e-3)


def test_emulated_kernels_random_configurations(emu):
    """Seeded sweep over odd sizes (images down to one pixel wide, R = 1..31, 1..700 splats, SH degree 0..3, coverage from
    sparse to 60x overdraw, scale_modifier != 1): forward and backward against the oracle. The sweep this sample is
    drawn from (100 configurations) found the texel-boundary conditioning case documented in DESIGN.md §7."""
    import random
    for it in (0, 9, 14, 24, 29, 31):
        rnd = random.Random(1000 + it)
        N = rnd.choice([1, 2, 7, 33, 100, 300, 700])
        W, H = rnd.randint(1, 80), rnd.randint(1, 60)
        R = rnd.choice([1, 2, 3, 5, 8, 16, 31])
        deg = rnd.randint(0, 3)
        cov = rnd.choice([1.0, 4.0, 16.0, 60.0])
        sm = rnd.choice([1.0, 1.0, 0.5, 1.7])
        bg = (rnd.random(), rnd.random(), rnd.random())
        g = sphere_shell_scene(N, R, sh_degree=deg, seed=it + 1, tex_seed=it + 1, coverage=cov)
        cam = orbit_cameras(1, W, H, seed=it + 2)[0]
        check_forward(g, cam, bg=bg, scale_modifier=sm, runner=run_emu, max_amb=1.0)
        check_backward(g, cam, bg=bg, scale_modifier=sm, runner=run_emu, max_flag=1.0, uv_tol=5e-3)


def test_emulated_kernels_survive_degenerate_inputs(emu):
    """NaN / inf positions, a zero quaternion, zero and huge scales, opacity 0 and 1, a zero uv vector and a zero
    Jacobian among ordinary splats: no hang, no emulator error, finite outputs and gradients, and the ordinary splats'
    pixels still match the oracle (the oracle's own gradients turn NaN for a NaN position; the kernels cull it)."""
    N, W, H, R = 200, 40, 30, 8
    g = sphere_shell_scene(N, R, sh_degree=2, seed=3, tex_seed=4, coverage=8.0)
    t = {k: (v.detach().clone() if v is not None else None) for k, v in g.tensors().items()}
    t["xyz"][5] = float("nan")
    t["xyz"][6] = float("inf")
    t["rotation"][7] = 0.0
    t["scaling"][9] = 0.0
    t["scaling"][11, :2] = 50.0
    t["opacity"][13] = 0.0
    t["opacity"][15] = 1.0
    t["uvs"][17] = 0.0
    t["grad_uvs"][19] = 0.0
    gg = SyntheticGaussians(active_sh_degree=2, **t)
    cam = orbit_cameras(1, W, H, seed=5)[0]
    ref, aux, _ = run_oracle(gg, cam, bg=(0.1, 0.2, 0.3))
    cot = [c * (~aux["grad_ambiguous"]).float() for c in output_cotangents(H, W, seed=3)]
    got, stats, grads = run_emu(gg, cam, bg=(0.1, 0.2, 0.3), cot=cot)
    assert all(bool(torch.isfinite(x).all()) for x in got[:4])
    assert all(bool(torch.isfinite(v).all()) for v in grads.values() if v is not None)
    assert int(got[4][5]) == 0 and int(got[4][6]) == 0 and float(grads["xyz"][5].abs().max()) == 0.0      # culled
    rep = compare_images(got[:4], ref[:4], aux["ambiguous"])
    for nme in ("image", "depth", "norm", "alpha"):
        assert rep[nme]["max_clear"] <= ABS_TOL * (3.0 if nme == "depth" else 1.0), (nme, rep[nme])
    assert int((got[4] != ref[4]).sum()) == 0


def test_reference_wrapper_source_drives_the_emulated_kernels(emu, monkeypatch):
    """Drop-in boundary, end to end on CPU: the reference's OWN ``render/uv_tex_render.py`` (executed where it lies; its
    one hard-coded ``device="cuda"`` re-targeted to the inputs' device) imports ``diff_gauss_uv_tex``, builds the 12-field
    settings and calls ``GaussianRasterizer(...)(means3D=..., ..., extra_attrs=None)``. Here that module is the emulated
    kernels behind the product's call signature; the returned dict matches the oracle and ``viewspace_points.grad`` is
    populated (what models/gaussian3d.py:335 reads). Skipped where /root/reference is absent (the GPU box)."""
    import inspect
    import sys
    import types
    from pathlib import Path
    src_path = Path("/root/reference/render/uv_tex_render.py")
    if not src_path.exists():
        pytest.skip("reference tree not present")
    from simt import emu_module
    from texture_gs_b200.rasterizer import GaussianRasterizer as ProductRasterizer
    assert inspect.signature(emu_module.GaussianRasterizer.forward) == inspect.signature(ProductRasterizer.forward)
    shim = types.ModuleType("diff_gauss_uv_tex")
    shim.GaussianRasterizationSettings = emu_module.GaussianRasterizationSettings
    shim.GaussianRasterizer = emu_module.GaussianRasterizer
    monkeypatch.setitem(sys.modules, "diff_gauss_uv_tex", shim)
    src = src_path.read_text()
    assert src.count('device="cuda"') == 1
    ns = {}
    exec(compile(src.replace('device="cuda"', "device=gaussians.get_xyz.device"), str(src_path), "exec"), ns)
    g = sphere_shell_scene(800, 16, sh_degree=3, seed=31, tex_seed=32).to("cpu", requires_grad=True)
    cam = orbit_cameras(1, 64, 48, seed=33)[0]
    bg = torch.tensor([0.2, 0.3, 0.1])
    pkg = ns["uv_tex_render"](cam, g, None, bg)
    assert set(pkg) == {"render", "depth", "norm", "alpha", "viewspace_points", "visibility_filter", "extra", "radii"}
    assert pkg["extra"] is None and pkg["visibility_filter"].dtype == torch.bool
    ref, aux, _ = run_oracle(g, cam, bg=(0.2, 0.3, 0.1))
    rep = compare_images([pkg[k].detach() for k in ("render", "depth", "norm", "alpha")], ref[:4], aux["ambiguous"])
    for nme in ("image", "depth", "norm", "alpha"):
        assert rep[nme]["max_clear"] <= ABS_TOL * (3.0 if nme == "depth" else 1.0), (nme, rep[nme])
    assert torch.equal(pkg["radii"], ref[4].to(pkg["radii"].dtype))
    keep = (~aux["grad_ambiguous"]).float()
    cot = [c * keep for c in output_cotangents(48, 64, seed=34)]
    sum((pkg[k] * c).sum() for k, c in zip(("render", "depth", "norm", "alpha"), cot)).backward()
    _, _, gref = run_oracle(g, cam, bg=(0.2, 0.3, 0.1), cot=cot)
    t = g.tensors()
    for k in ("xyz", "opacity", "scaling", "rotation", "shs", "uvs", "texture"):
        e = rel_err(t[k].grad, gref[k])
        assert e <= (5e-3 if k == "uvs" else GRAD_RTOL), (k, e)
    vg = pkg["viewspace_points"].grad
    assert vg is not None and rel_err(vg[:, :2], gref["means2D"][:, :2]) <= GRAD_RTOL and float(vg[:, 2].abs().max()) == 0.0


def test_emulated_list_longer_than_the_shared_memory_sort_takes_the_global_network(emu):
    """> 4096 splats in one tile: the sort kernel's global-memory bitonic network with virtual +inf padding."""
    n = 4300
    g = sphere_shell_scene(n, 8, sh_degree=0, seed=5, coverage=4.0)
    t = g.tensors()
    cam = orbit_cameras(1, 16, 16, seed=6)[0]
    c = cam.camera_center / cam.camera_center.norm()
    xyz = c[None, :] * 1.0 + 0.01 * torch.randn(n, 3, generator=torch.Generator().manual_seed(0))
    gg = SyntheticGaussians(active_sh_degree=0, **{**{k: (v.detach() if v is not None else None) for k, v in t.items()},
                                                    "xyz": xyz, "uvs": xyz / xyz.norm(dim=1, keepdim=True), "opacity": torch.full((n, 1), 0.02)})
    ref, aux, _ = run_oracle(gg, cam)
    got, stats, _ = run_emu(gg, cam)
    assert stats.max_tile_len > 4096 and float(got[3].max()) > 0.5
    res = run_emu.last
    order = res.sorted_ids.numpy().astype(np.int64)
    assert sorted(order.tolist()) == list(range(n))                              # a permutation: nothing lost in the padding
    d = aux["pre"]["depth"].detach().numpy()
    # depth order; 4300 depths inside 0.08 units leave a handful of pairs within an ulp of each other, which the kernel
    # (its own rounding of z) may order the other way round than the oracle (see the depth-tie flag of the oracle)
    assert (np.diff(d[order]) >= -4e-7).all()
    rep = compare_images(got[:4], ref[:4], aux["ambiguous"])
    for nme in ("image", "alpha"):
        assert rep[nme]["max_clear"] <= 2 * ABS_TOL, rep


def test_comparison_against_the_c_oracle_used_at_full_size_on_the_gpu(emu):
    """tests/test_gpu_zz_fullsize.py compares the CUDA kernels with the C oracle at 500 k / 1080p; the same helper is run
    here with the emulated kernels on a 1/25-scale copy of that scene (same generator, same splats per pixel); the full-size
    run takes three minutes on CPU (tools/emu_check.py, profiles/r1_emulated_kernels_vs_c_oracle.md)."""
    from util import check_against_c_oracle
    g = sphere_shell_scene(20_000, 512, sh_degree=3, seed=0)
    cam = orbit_cameras(32, 384, 216, seed=1)[5]
    emu.build().simt_set_fastmath_noise(3)          # the hardware's approximate units: ~2 ulp, repeatable
    try:
        check_against_c_oracle(g, cam, bg=(0.1, 0.2, 0.3), runner=run_emu, max_flag=0.6)
    finally:
        emu.build().simt_set_fastmath_noise(0)


def test_tuning_macro_variants_compute_the_same_thing(emu):
    """The build-time tuning knobs used for A/B timing on the GPU (tools/build_variants.py) must not change results:
    16-entry chunks (TEXGS_CHUNK), the unstaged SH path of preprocess_bwd (TEXGS_PREBWD_STAGE_SH=0) and libm exp instead
    of the fast intrinsic (TEXGS_FAST_EXP=0), all in one alternative build that lives in the same process as the default
    one. Forward bit-identical (the blend order does not depend on the chunking), gradients equal up to summation order."""
    alt = emu.build(extra_flags=("-DTEXGS_CHUNK=16", "-DTEXGS_PREBWD_STAGE_SH=0", "-DTEXGS_FAST_EXP=0"))
    g = sphere_shell_scene(1200, 32, sh_degree=3, seed=5, tex_seed=6)
    cam = orbit_cameras(1, 96, 64, seed=7)[0]
    t = g.tensors()
    cot = output_cotangents(64, 96, seed=8)
    kw = dict(means3D=t["xyz"], opacities=t["opacity"], scales=t["scaling"], rotations=t["rotation"], shs=t["shs"], uvs=t["uvs"],
              gradient_uvs=t["grad_uvs"], texture=t["texture"], cotangents=cot, **_cam_kw(cam, (0.1, 0.2, 0.3), 3))
    a = emu.rasterize(**kw)
    b = emu.rasterize(lib=alt, **kw)
    for x, y in zip((a.image, a.depth, a.norm, a.alpha, a.radii), (b.image, b.depth, b.norm, b.alpha, b.radii)):
        assert torch.equal(x, y)
    for k in a.grads:
        assert rel_err(b.grads[k], a.grads[k]) < 1e-5, k


def test_full_size_criterion_is_not_vacuous(emu):
    """Negative controls for tests/util.check_against_c_oracle (the comparison tests/test_gpu_zz_fullsize.py runs): a
    renderer whose image is off by 3e-4 on 0.1 % of the pixels, one whose uv gradient is 1 % too large, and one that
    drops a splat are all rejected; the unmodified emulated kernels pass."""
    from util import check_against_c_oracle
    g = sphere_shell_scene(3000, 64, sh_degree=3, seed=0)
    cam = orbit_cameras(32, 160, 90, seed=1)[5]

    def broken(kind):
        def runner(gg, cc, bg=(0, 0, 0), cot=None, scale_modifier=1.0):
            if kind == "drop":
                t = {k: (v.detach().clone() if v is not None else None) for k, v in gg.tensors().items()}
                t["opacity"][::50] = 0.0
                gg = SyntheticGaussians(active_sh_degree=gg.active_sh_degree, **t)
            outs, stats, grads = run_emu(gg, cc, bg=bg, cot=cot, scale_modifier=scale_modifier)
            if kind == "image":
                img = outs[0].clone()
                img.view(3, -1)[:, ::1000] += 3e-4
                outs = (img,) + tuple(outs[1:])
            if kind == "uvs" and grads is not None:
                grads = dict(grads, uvs=grads["uvs"] * 1.01)
            return outs, stats, grads
        return runner

    check_against_c_oracle(g, cam, bg=(0.1, 0.2, 0.3), runner=run_emu, max_flag=0.6)
    for kind in ("image", "uvs", "drop"):
        with pytest.raises(AssertionError):
            check_against_c_oracle(g, cam, bg=(0.1, 0.2, 0.3), runner=broken(kind), max_flag=0.6)


def test_reference_plain_render_source_and_its_own_eval_sh_drive_the_emulated_kernels(emu, monkeypatch):
    """The diff_gauss side of the boundary: the reference's OWN ``render/render.py`` (``device="cuda"`` re-targeted) with
    its OWN ``utils/sh.py`` runs against the emulated kernels. ``cfg.convert_SHs_python = True`` makes the reference
    evaluate the spherical harmonics itself (``eval_sh``) and hand over ``colors_precomp``; ``False`` hands the
    coefficients to the rasterizer. Both must give the same picture — which pins the kernels' SH evaluation to the
    reference's own code — and ``compute_cov3D_python`` (``cov3Ds_precomp``) must agree with scales + rotations."""
    import sys
    import types
    from pathlib import Path
    ref = Path("/root/reference")
    if not (ref / "render" / "render.py").exists():
        pytest.skip("reference tree not present")
    from simt import emu_module
    for name in ("diff_gauss",):
        shim = types.ModuleType(name)
        shim.GaussianRasterizationSettings = emu_module.GaussianRasterizationSettings
        shim.GaussianRasterizer = emu_module.GaussianRasterizer
        monkeypatch.setitem(sys.modules, name, shim)
    sh_mod = types.ModuleType("utils.sh")
    exec(compile((ref / "utils" / "sh.py").read_text(), str(ref / "utils" / "sh.py"), "exec"), sh_mod.__dict__)
    pkg = types.ModuleType("utils")
    pkg.sh = sh_mod
    monkeypatch.setitem(sys.modules, "utils", pkg)
    monkeypatch.setitem(sys.modules, "utils.sh", sh_mod)
    src = (ref / "render" / "render.py").read_text()
    assert src.count('device="cuda"') == 1
    ns = {}
    exec(compile(src.replace('device="cuda"', "device=gaussians.get_xyz.device"), str(ref / "render" / "render.py"), "exec"), ns)

    N, W, H = 700, 64, 48
    g0 = sphere_shell_scene(N, 4, sh_degree=3, seed=61)
    t = g0.tensors()
    gen = torch.Generator().manual_seed(62)
    feats = torch.cat([torch.randn(N, 1, 3, generator=gen), 0.1 * torch.randn(N, 15, 3, generator=gen)], dim=1)

    class G:      # the attributes render/render.py reads (models/gaussian3d.py)
        get_xyz, get_opacity, get_scaling, get_rotation = t["xyz"].detach(), t["opacity"].detach(), t["scaling"].detach(), t["rotation"].detach()
        get_features, max_sh_degree, active_sh_degree = feats, 3, 3

        @staticmethod
        def get_covariance(scaling_modifier=1.0):
            from oracle.raster_ref import quat_to_rot
            Lm = quat_to_rot(G.get_rotation.double()) * (G.get_scaling.double() * scaling_modifier)[:, None, :]
            S = Lm @ Lm.transpose(1, 2)
            return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=-1).float()

    cam = orbit_cameras(1, W, H, seed=63)[0]
    bg = torch.tensor([0.2, 0.1, 0.3])
    Cfg = lambda sh_py, cov_py: types.SimpleNamespace(convert_SHs_python=sh_py, compute_cov3D_python=cov_py)
    a = ns["render"](cam, G, Cfg(False, False), bg)          # SH evaluated by the kernels
    b = ns["render"](cam, G, Cfg(True, False), bg)           # SH evaluated by the reference's eval_sh
    c = ns["render"](cam, G, Cfg(False, True), bg)           # covariance computed "in Python"
    assert set(a) == {"render", "depth", "norm", "alpha", "viewspace_points", "visibility_filter", "extra", "radii"}
    assert float((a["render"] - b["render"]).detach().abs().max()) <= 2e-6, "kernel SH evaluation differs from the reference's eval_sh"
    for k in ("depth", "alpha", "norm"):
        assert torch.equal(a[k], b[k])
    # covariance path: same picture on well-conditioned pixels (the two covariances differ by rounding)
    from oracle import raster_ref as RR
    st = oracle_settings(cam, 3, bg=(0.2, 0.1, 0.3))
    o = RR.rasterize(G.get_xyz, None, feats, G.get_opacity, G.get_scaling, G.get_rotation, None, None, None, st, return_aux=True)
    amb = o[-1]["ambiguous"]
    assert float((a["render"] - o[0]).detach().abs().amax(0)[~amb].max()) <= ABS_TOL
    assert float((c["render"] - a["render"]).detach().abs().amax(0)[~amb].max()) <= ABS_TOL
    assert torch.equal(a["radii"], o[4].to(a["radii"].dtype))


def test_textured_sh_rest_path_equals_the_full_sh_path_for_a_constant_texture(emu):
    """The textured mode evaluates bands l >= 1 from ``shs`` (N,15,3) and takes the DC term from the texture
    (models/texture_gaussian3d.py:97-98); the diff_gauss mode evaluates all 16 coefficients — the path pinned to the
    reference's ``eval_sh`` above. With a constant texture t0 the two must give the same image when the full SH set is
    [t0, rest]: the pin carries over to the SH-rest evaluation of the textured kernels."""
    N, W, H = 900, 64, 48
    g = sphere_shell_scene(N, 4, sh_degree=3, seed=71)
    t = g.tensors()
    cam = orbit_cameras(1, W, H, seed=72)[0]
    t0 = torch.tensor([0.7, -0.4, 1.1])
    tex = t0.expand(6, 4, 4, 3).contiguous()
    common = dict(means3D=t["xyz"], opacities=t["opacity"], scales=t["scaling"], rotations=t["rotation"], **_cam_kw(cam, (0.1, 0.2, 0.3), 3))
    a = emu.rasterize(shs=t["shs"], uvs=t["uvs"], gradient_uvs=t["grad_uvs"], texture=tex, **common)
    full = torch.cat([t0.expand(N, 1, 3), t["shs"].detach()], dim=1).contiguous()
    b = emu.rasterize(shs=full, **common)
    assert float((a.image - b.image).abs().max()) <= 2e-6
    for x, y in zip((a.depth, a.norm, a.alpha, a.radii), (b.depth, b.norm, b.alpha, b.radii)):
        assert torch.equal(x, y)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_emulated_kernels_under_random_warp_interleavings(emu, seed):
    """The warps of a block in random order, with random time slices and a random first lane (the default schedule runs
    warp after warp, each as far as it can go): the sort kernels' barriers, the scan, the per-warp rings of the render
    kernels and the staged SH rows of preprocess_bwd must give the same results under any interleaving — forward
    bit-identical, gradients up to the order of float additions."""
    g = sphere_shell_scene(1200, 32, sh_degree=3, seed=5, tex_seed=6)
    cam = orbit_cameras(1, 96, 64, seed=7)[0]
    t = g.tensors()
    cot = output_cotangents(64, 96, seed=8)
    kw = dict(means3D=t["xyz"], opacities=t["opacity"], scales=t["scaling"], rotations=t["rotation"], shs=t["shs"], uvs=t["uvs"],
              gradient_uvs=t["grad_uvs"], texture=t["texture"], cotangents=cot, **_cam_kw(cam, (0.1, 0.2, 0.3), 3))
    lib = emu.build()
    a = emu.rasterize(**kw)
    lib.simt_set_schedule_seed(seed)
    lib.simt_set_eager_copies(seed & 1)
    try:
        b = emu.rasterize(**kw)
    finally:
        lib.simt_set_schedule_seed(0)
        lib.simt_set_eager_copies(0)
    for x, y in zip((a.image, a.depth, a.norm, a.alpha, a.radii, a.sorted_ids), (b.image, b.depth, b.norm, b.alpha, b.radii, b.sorted_ids)):
        assert torch.equal(x, y)
    for k in a.grads:
        assert rel_err(b.grads[k], a.grads[k]) < 1e-5, k


@pytest.mark.parametrize("name", ["seamless_cube", "depth_of_intersection", "stopgrad_delta", "all"])
def test_emulated_spec_switch_instantiations_match_the_oracle_with_the_same_switch(emu, name):
    """the conventions the reference tree does not pin are run-time flags of the
    library (TEXGS_FLAG_SEAMLESS_CUBE / DEPTH_INTERSECTION / STOPGRAD_DELTA -> cold ALT instantiations of the render
    kernels); each must reproduce the oracle run with the same switch, forward and backward, and must actually change
    the result against the default (a switch that does nothing would pass the parity test vacuously). Low texture
    resolution on purpose: many bilinear footprints straddle a face edge."""
    from oracle.raster_ref import Switches
    sw = Switches(True, True, True) if name == "all" else Switches(**{name: True})
    g = sphere_shell_scene(1200, 8, sh_degree=2, seed=21, tex_seed=22)
    cam = orbit_cameras(1, 96, 64, seed=23)[0]
    check_forward(g, cam, bg=(0.1, 0.3, 0.2), runner=run_emu, max_amb=0.3, sw=sw)
    errs = check_backward(g, cam, bg=(0.1, 0.3, 0.2), runner=run_emu, uv_tol=5e-3, max_flag=0.35, sw=sw)
    assert set(errs) >= {"xyz", "rotation", "uvs", "texture"}
    # non-vacuous: the default convention gives a different picture / different gradients
    cot = output_cotangents(64, 96, seed=3)
    a, _, ga = run_emu(g, cam, bg=(0.1, 0.3, 0.2), cot=cot)
    b, _, gb = run_emu(g, cam, bg=(0.1, 0.3, 0.2), cot=cot, sw=sw)
    if sw.seamless_cube:
        assert float((a[0] - b[0]).abs().max()) > 1e-3 and rel_err(gb["texture"], ga["texture"]) > 1e-3
    if sw.depth_of_intersection:
        assert float((a[1] - b[1]).abs().max()) > 1e-4
    if sw.stopgrad_delta and not sw.depth_of_intersection:
        assert torch.equal(a[0], b[0]) and rel_err(gb["xyz"], ga["xyz"]) > 1e-3 and rel_err(gb["uvs"], ga["uvs"]) < 1e-6


def test_spec_switches_need_the_packed_texel_paths(emu):
    """The ALT instantiations exist for the packed texel copy / padded texel gradient only; the plain layouts refuse."""
    g = sphere_shell_scene(50, 8, sh_degree=0, seed=1)
    cam = orbit_cameras(1, 32, 32, seed=2)[0]
    t = g.tensors()
    with pytest.raises(Exception, match="packed texel copy"):
        emu.rasterize(means3D=t["xyz"], opacities=t["opacity"], scales=t["scaling"], rotations=t["rotation"], shs=t["shs"], uvs=t["uvs"],
                      gradient_uvs=t["grad_uvs"], texture=t["texture"], packed_texture=False, spec_flags=4, **_cam_kw(cam, (0, 0, 0), 0))
