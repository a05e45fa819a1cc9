"""The SIMT emulator itself (tests/simt/simt_emu.h) and the non-rasterizer kernels under it.

1. ``selftest.cpp``: kernels with known answers, and kernels with known BUGS that the emulator must report (divergent
   barrier, collective after a lane exited, shared memory read before the mbarrier wait, block exit with a bulk copy
   in flight, wrong expect_tx byte count, stage overwritten while being read) — an emulator that ran everything
   "successfully" would prove nothing about the kernels it is used on.
2. The fused loss kernels against the oracle pinned by the reference's own ``losses/*.py``, and the
   fused texture Adam kernel (N4) against ``torch.optim.Adam`` (the reference's optimizer), through the C-ABI of the
   emulated library.
"""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import loss_ref

HERE = Path(__file__).resolve().parent
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the SIMT emulator needs g++")


def test_emulator_selftest_known_answers_and_known_bugs(tmp_path):
    exe = tmp_path / "selftest"
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fno-extern-tls-init", "-Wno-attributes", "-Wno-unknown-pragmas",
                    "-I", str(HERE / "simt"), str(HERE / "simt" / "selftest.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "SELFTEST PASSED" in r.stdout and "FAIL" not in r.stdout
    assert r.stdout.count("ok ") >= 11


@pytest.fixture(scope="module", params=[0, 5], ids=["warp-after-warp", "random-interleaving"])
def lib(request):
    """The emulated library, once with the deterministic schedule and once with the warps of every block interleaved at
    random (block barriers and shared-memory staging of the loss / optimizer kernels must not depend on the order)."""
    from simt import emu
    lb = emu.build()
    lb.simt_set_schedule_seed(request.param)
    yield lb
    lb.simt_set_schedule_seed(0)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _ws(nbytes):
    from simt.emu import _aligned_empty
    w = _aligned_empty(max(nbytes, 256))
    w.fill_(0xA5)
    return w


@pytest.mark.parametrize("H,W", [(37, 53), (16, 16)])
def test_emulated_photometric_loss_kernels_match_reference_pinned_oracle(lib, H, W):
    """(1-l)*L1 + l*(1-SSIM) of models/texture_gaussian3d.py:333-340, forward values and dL/dimage."""
    from simt.emu import check
    gen = torch.Generator().manual_seed(H)
    img = torch.rand(3, H, W, generator=gen).requires_grad_(True)
    gt = (img.detach() + 0.2 * torch.randn(3, H, W, generator=gen)).clamp(0, 1)
    lam = 0.2
    loss, l1, lssim = loss_ref.photometric_loss(img.double(), gt.double(), lam)
    coef = torch.tensor([0.7, -0.4, 1.3])            # cotangents of (loss, Ll1, Lssim)
    (coef[0] * loss + coef[1] * l1 + coef[2] * lssim).backward()
    n = C.c_size_t()
    check(lib, lib.texgs_photometric_workspace_size(3, H, W, C.byref(n)), "ws")
    ws, out3 = _ws(n.value), torch.empty(3)
    x = img.detach().contiguous()
    check(lib, lib.texgs_photometric_forward(_p(x), _p(gt), 3, H, W, lam, _p(ws), _p(out3), None), "photometric_forward")
    for got, ref in zip(out3.tolist(), (loss.item(), l1.item(), lssim.item())):
        assert abs(got - ref) <= 2e-6 + 1e-5 * abs(ref)
    coef2 = torch.tensor([coef[0] * (1 - lam) + coef[1], coef[0] * lam + coef[2]])
    dimg = torch.full_like(x, float("nan"))
    check(lib, lib.texgs_photometric_backward(_p(x), _p(gt), 3, H, W, _p(ws), _p(coef2), _p(dimg), None), "photometric_backward")
    ref = img.grad
    assert float((dimg.double() - ref.double()).abs().max()) <= 1e-3 * float(ref.abs().max())


def test_emulated_geometry_loss_kernels_match_reference_pinned_oracle(lib):
    """Lalpha / Lnorm / Lnsm of models/texture_gaussian3d.py:342-368 (losses/norm_reg_loss.py:66-71, smooth_loss.py:4-27)."""
    from simt.emu import check
    H, W = 29, 70            # three 32x8 tiles wide with a ragged edge, four tiles high
    gen = torch.Generator().manual_seed(3)
    alpha = torch.rand(1, H, W, generator=gen).requires_grad_(True)
    norm = torch.randn(3, H, W, generator=gen).requires_grad_(True)
    gt_alpha = (torch.rand(1, H, W, generator=gen) > 0.3).float()
    gt_norm = torch.nn.functional.normalize(torch.randn(3, H, W, generator=gen), dim=0)
    gt_image = torch.rand(3, H, W, generator=gen)
    la, ln, ls = loss_ref.geometry_losses(alpha.double(), norm.double(), gt_alpha.double(), gt_norm.double(), gt_image.double(), 0.1)
    coef = torch.tensor([0.5, -1.5, 2.0])
    (coef[0] * la + coef[1] * ln + coef[2] * ls).backward()
    n = C.c_size_t()
    check(lib, lib.texgs_geometry_loss_workspace_size(H, W, C.byref(n)), "ws")
    ws, out3 = _ws(n.value), torch.empty(3)
    a, nm = alpha.detach().contiguous(), norm.detach().contiguous()
    check(lib, lib.texgs_geometry_loss_forward(_p(a), _p(nm), _p(gt_alpha), _p(gt_norm), _p(gt_image), H, W, 0.1, _p(ws), _p(out3), None), "geo_fwd")
    for got, ref in zip(out3.tolist(), (la.item(), ln.item(), ls.item())):
        assert abs(got - ref) <= 2e-6 + 1e-5 * abs(ref), (out3, la, ln, ls)
    d_a, d_n = torch.full_like(a, float("nan")), torch.full_like(nm, float("nan"))
    check(lib, lib.texgs_geometry_loss_backward(_p(a), _p(nm), _p(gt_alpha), _p(gt_norm), _p(gt_image), H, W, 0.1, _p(ws), _p(coef),
                                                _p(d_a), _p(d_n), None), "geo_bwd")
    assert float((d_a.double() - alpha.grad.double()).abs().max()) <= 1e-4 * float(alpha.grad.abs().max()) + 1e-9
    assert float((d_n.double() - norm.grad.double()).abs().max()) <= 1e-3 * float(norm.grad.abs().max())


@pytest.mark.parametrize("padded", [True, False])
def test_emulated_texture_adam_kernel_matches_torch_adam(lib, padded):
    """Three steps of the fused texture optimizer kernel == torch.optim.Adam(eps=1e-15) (models/texture_gaussian3d.py:139-143)
    on a (6,R,R,3) texture whose texel count is not a multiple of the kernel's 1024-texel CTA tile; the padded
    gradient is cleared in the same pass and the packed RGBA copy written."""
    from simt.emu import _aligned_empty, check
    R = 19
    ntex = 6 * R * R
    gen = torch.Generator().manual_seed(7)
    p_ref = torch.randn(6, R, R, 3, generator=gen).requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=2.5e-3, betas=(0.9, 0.999), eps=1e-15)
    p = _aligned_empty(ntex * 12).view(torch.float32).view(6, R, R, 3)
    p.copy_(p_ref.detach())
    m, v = _aligned_empty(ntex * 12).view(torch.float32), _aligned_empty(ntex * 12).view(torch.float32)
    m.zero_(); v.zero_()
    rgba = _aligned_empty(ntex * 16).view(torch.float32).view(6, R, R, 4)
    rgba.fill_(float("nan"))
    for step in range(1, 4):
        g = torch.randn(6, R, R, 3, generator=gen) * (10.0 ** (step - 3))
        g[0, :3] = 0.0                                        # untouched texels keep moving through their moments
        p_ref.grad = g.clone()
        opt.step()
        if padded:
            g4 = _aligned_empty(ntex * 16).view(torch.float32).view(6, R, R, 4)
            g4[..., :3] = g
            g4[..., 3] = 0.0
            check(lib, lib.texgs_texture_adam_step(_p(p), _p(m), _p(v), None, _p(g4), _p(rgba), ntex, 2.5e-3, 0.9, 0.999, 1e-15, step, 1, None), "adam")
            assert float(g4.abs().max()) == 0.0               # cleared for the next accumulation
        else:
            g3 = _aligned_empty(ntex * 12).view(torch.float32).view(6, R, R, 3)
            g3.copy_(g)
            check(lib, lib.texgs_texture_adam_step(_p(p), _p(m), _p(v), _p(g3), None, _p(rgba), ntex, 2.5e-3, 0.9, 0.999, 1e-15, step, 0, None), "adam")
            assert torch.equal(g3, g)
        assert float((p - p_ref.detach()).abs().max()) <= 2e-6, step
        assert torch.equal(rgba[..., :3], p) and float(rgba[..., 3].abs().max()) == 0.0
    st = opt.state[p_ref]
    np.testing.assert_allclose(m.view(6, R, R, 3).numpy(), st["exp_avg"].numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(v.view(6, R, R, 3).numpy(), st["exp_avg_sq"].numpy(), rtol=1e-5, atol=1e-12)


def test_emulated_mark_visible_matches_the_near_plane_rule(lib):
    """GaussianRasterizer.markVisible (upstream API, unused by the reference tree): z_view > 0.2 (spec E1)."""
    from simt.emu import check
    from texture_gs_b200.scene import orbit_cameras
    cam = orbit_cameras(1, 64, 48, seed=3)[0]
    pos = torch.randn(777, 3, generator=torch.Generator().manual_seed(1)) * 2.0
    present = torch.full((777,), -1, dtype=torch.int32)
    vm = (C.c_float * 16)(*cam.world_view_transform.reshape(-1).tolist())
    pm = (C.c_float * 16)(*cam.full_proj_transform.reshape(-1).tolist())
    check(lib, lib.texgs_mark_visible(777, _p(pos), vm, pm, _p(present), None), "texgs_mark_visible")
    z = pos @ cam.world_view_transform[:3, 2] + cam.world_view_transform[3, 2]
    clear = (z - 0.2).abs() > 1e-5
    assert torch.equal(present.bool()[clear], (z > 0.2)[clear]) and int(present.min()) >= 0


@pytest.mark.parametrize("world", [2, 3])
def test_emulated_data_parallel_texture_step_equals_summed_gradient_adam(lib, world):
    """The fused multi-GPU texture step (texgs_texture_adam_dp_step, peer path) with the ranks emulated as buffers of one
    process: every "rank" holds a partial padded gradient and a copy of the texture; after each rank has run its launch
    (it owns 1/world of the 1024-texel tiles, pulls and adds all partial gradients of those, updates its shard of the
    moments, pushes the texels into every copy) all copies equal torch.optim.Adam(eps=1e-15) applied to the SUMMED gradient,
    the gradients are untouched, and the concatenated moment shards are Adam's moments. Texel count not a multiple of the
    tile (the tail tile belongs to the last rank), three steps."""
    from simt.emu import _aligned_empty, check
    from texture_gs_b200 import _lib as L
    R = 23
    ntex = 6 * R * R                                    # 3174 texels: 3 full tiles + a tail
    gen = torch.Generator().manual_seed(11)
    p_ref = torch.randn(6, R, R, 3, generator=gen).requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=2.5e-3, betas=(0.9, 0.999), eps=1e-15)
    params = [_aligned_empty(ntex * 12).view(torch.float32) for _ in range(world)]
    grads = [_aligned_empty(ntex * 16).view(torch.float32).view(ntex, 4) for _ in range(world)]
    for p in params:
        p.copy_(p_ref.detach().reshape(-1))
    shards = []
    for r in range(world):
        lo, hi = C.c_uint64(), C.c_uint64()
        check(lib, lib.texgs_dp_shard(ntex, world, r, C.byref(lo), C.byref(hi)), "shard")
        n = max(1, (hi.value - lo.value) * 1024 * 3)
        m, v = _aligned_empty(n * 4).view(torch.float32), _aligned_empty(n * 4).view(torch.float32)
        m.zero_(); v.zero_()
        shards.append((lo.value, hi.value, m, v))
    assert shards[0][0] == 0 and shards[-1][1] == (ntex + 1023) // 1024 and all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
    for step in range(1, 4):
        parts = [torch.randn(ntex, 3, generator=gen) * (10.0 ** (step - 3)) for _ in range(world)]
        for gbuf, part in zip(grads, parts):
            gbuf[:, :3] = part
            gbuf[:, 3] = 7.0                              # the pad float of a texel never reaches the update
        p_ref.grad = sum(parts).reshape(6, R, R, 3).clone()
        opt.step()
        for r in range(world):
            a = L.TexgsDpAdamArgs()
            a.world, a.rank = world, r
            for k in range(world):
                a.grad_ptrs[k], a.param_ptrs[k] = grads[k].data_ptr(), params[k].data_ptr()
            lo, hi, m, v = shards[r]
            a.exp_avg, a.exp_avg_sq = m.data_ptr(), v.data_ptr()
            a.n_texels, a.tile_lo, a.tile_hi = ntex, lo, hi
            a.lr, a.beta1, a.beta2, a.eps, a.step = 2.5e-3, 0.9, 0.999, 1e-15, step
            check(lib, lib.texgs_texture_adam_dp_step(C.byref(a), None), "dp_step")
        for k in range(world):
            assert float((params[k] - p_ref.detach().reshape(-1)).abs().max()) <= 2e-6, (step, k)
            assert torch.equal(params[k], params[0])
            assert torch.equal(grads[k][:, :3], parts[k]) and float((grads[k][:, 3] - 7.0).abs().max()) == 0.0
    st = opt.state[p_ref]
    for lo, hi, m, v in shards:
        t0, t1 = lo * 1024, min(hi * 1024, ntex)
        # the kernel adds the partial gradients rank by rank starting at its own: a different order than sum(parts)
        np.testing.assert_allclose(m[:(t1 - t0) * 3].numpy(), st["exp_avg"].reshape(-1)[t0 * 3:t1 * 3].numpy(), rtol=2e-4, atol=1e-8)
        np.testing.assert_allclose(v[:(t1 - t0) * 3].numpy(), st["exp_avg_sq"].reshape(-1)[t0 * 3:t1 * 3].numpy(), rtol=2e-4, atol=1e-11)
    # the multicast variant needs the fabric: refused on the host
    a.grad_mc, a.param_mc = grads[0].data_ptr(), params[0].data_ptr()
    assert lib.texgs_texture_adam_dp_step(C.byref(a), None) != 0
