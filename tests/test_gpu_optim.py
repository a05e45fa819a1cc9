"""GPU parity of the fused texture optimizer step against torch.optim.Adam — the optimizer the
reference itself uses for the texture (models/texture_gaussian3d.py:139-143, 439-440)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected (-m gpu) but no CUDA device is visible")
    from texture_gs_b200 import _lib
    _lib.load()


def _grads(shape, steps, seed, sparse=0.5):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        t = torch.randn(shape, generator=g) * 10 ** torch.empty(shape).uniform_(-6, 0, generator=g)
        out.append((t * (torch.rand(shape, generator=g) > sparse)).cuda())       # many exact zeros, as untouched texels have
    return out


@pytest.mark.parametrize("R", [16, 13, 64])        # 1536 texels (one full tile + tail), 1014 (tail only), 24576 (full tiles)
def test_matches_torch_adam_over_several_steps(R):
    from texture_gs_b200.optim import TextureAdam
    torch.manual_seed(R)
    init = torch.randn(6, R, R, 3, device="cuda")
    p_ref, p_new = torch.nn.Parameter(init.clone()), torch.nn.Parameter(init.clone())
    ref = torch.optim.Adam([{"params": [p_ref], "lr": 0.0025}], lr=0.0, eps=1e-15)
    new = TextureAdam([{"params": [p_new], "lr": 0.0025}], lr=0.0, eps=1e-15)
    for i, g in enumerate(_grads(init.shape, 6, R)):
        p_ref.grad, p_new.grad = g.clone(), g.clone()
        ref.step(); new.step()
        if i == 2:
            for opt in (ref, new):
                opt.param_groups[0]["lr"] = 0.001                      # lr schedules edit param_groups in place
        assert float((p_new - p_ref).abs().max()) <= 2e-6 * float(p_ref.abs().max()), i
    sr, sn = ref.state[p_ref], new.state[p_new]
    assert float(sn["step"]) == float(sr["step"]) == 6.0
    assert float((sn["exp_avg"] - sr["exp_avg"]).abs().max()) <= 2e-6 * float(sr["exp_avg"].abs().max())
    assert float((sn["exp_avg_sq"] - sr["exp_avg_sq"]).abs().max()) <= 2e-6 * float(sr["exp_avg_sq"].abs().max())


def test_state_dicts_interchange_with_torch_adam():
    from texture_gs_b200.optim import TextureAdam
    init = torch.randn(6, 8, 8, 3, device="cuda")
    gs = _grads(init.shape, 4, 3)
    pa, pb = torch.nn.Parameter(init.clone()), torch.nn.Parameter(init.clone())
    a, b = torch.optim.Adam([pa], lr=0.01, eps=1e-15), TextureAdam([pb], lr=0.01, eps=1e-15)
    for g in gs[:2]:
        pa.grad, pb.grad = g.clone(), g.clone()
        a.step(); b.step()
    # swap: torch continues from the fused optimizer's state and vice versa (reference checkpoints: :150-155, :191-193)
    pa2, pb2 = torch.nn.Parameter(pb.detach().clone()), torch.nn.Parameter(pa.detach().clone())
    a2, b2 = torch.optim.Adam([pa2], lr=0.01, eps=1e-15), TextureAdam([pb2], lr=0.01, eps=1e-15)
    a2.load_state_dict(b.state_dict()); b2.load_state_dict(a.state_dict())
    for g in gs[2:]:
        pa2.grad, pb2.grad = g.clone(), g.clone()
        a2.step(); b2.step()
    assert float((pa2 - pb2).abs().max()) <= 2e-6 * float(pa2.abs().max())


def test_padded_bucket_gradient_zeroing_and_packed_copy_feed_the_next_render():
    """End to end on the path: render -> backward into a GradBucket (padded texture gradient) -> TextureAdam step that
    consumes the padded buffer, clears it and emits the packed texture -> next render equals a render that repacks."""
    from texture_gs_b200 import uv_tex_render, rasterizer as RZ
    from texture_gs_b200.dist import GradBucket
    from texture_gs_b200.optim import TextureAdam
    from texture_gs_b200.scene import sphere_shell_scene, orbit_cameras
    g = sphere_shell_scene(3000, 32, device="cuda")
    cam = orbit_cameras(1, 96, 64, device="cuda")[0]
    bg = torch.zeros(3, device="cuda")
    tex = g.get_texture
    assert tex.requires_grad
    bucket = GradBucket({"texture": tex})
    opt = TextureAdam([tex], lr=0.01, eps=1e-15, zero_grad_in_step=True)
    ref_p = torch.nn.Parameter(tex.detach().clone())
    ref = torch.optim.Adam([ref_p], lr=0.01, eps=1e-15)
    for _ in range(2):
        with bucket.fused():
            uv_tex_render(cam, g, None, bg)["render"].square().sum().backward()
        assert float(tex.grad.abs().max()) > 0
        ref_p.grad = tex.grad.clone().contiguous()
        ref.step(); opt.step()
        assert float(bucket.flat.abs().max()) == 0.0                  # cleared in the same pass
        assert float((tex - ref_p).abs().max()) <= 2e-6 * float(ref_p.abs().max())
        with torch.no_grad():
            img_cached = uv_tex_render(cam, g, None, bg)["render"].clone()    # uses the packed copy the optimizer emitted
            RZ.invalidate_packed_cache()
            img_fresh = uv_tex_render(cam, g, None, bg)["render"]             # repacks from the (6,R,R,3) texture
        assert torch.equal(img_cached, img_fresh)


def test_full_size_step_time_and_bandwidth():
    """R = 2048 (302 MB texture): fused step vs torch.optim.Adam + zero fill + repack, prints both and the achieved GB/s."""
    from texture_gs_b200.optim import TextureAdam
    from texture_gs_b200 import rasterizer as RZ, _lib as L
    import ctypes as C
    R = 2048
    p_ref = torch.nn.Parameter(torch.randn(6, R, R, 3, device="cuda"))
    p_new = torch.nn.Parameter(p_ref.detach().clone())
    pad = torch.randn(6, R, R, 4, device="cuda")
    p_new.grad = pad[..., :3]
    ref, new = torch.optim.Adam([p_ref], lr=0.0025, eps=1e-15), TextureAdam([p_new], lr=0.0025, eps=1e-15, zero_grad_in_step=True)
    tex4 = torch.empty(6, R, R, 4, device="cuda")
    lib = L.load()

    def torch_step():
        p_ref.grad = torch.zeros_like(p_ref)          # what zero_grad(set_to_none=True) + the next backward's fill cost
        ref.step()
        L.check(lib.texgs_pack_texture(C.c_void_p(p_ref.data_ptr()), R, C.c_void_p(tex4.data_ptr()),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pack")

    def timeit(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_ref, t_new = timeit(torch_step), timeit(new.step)
    gbs = 6 * R * R * 120 / (t_new * 1e-3) / 1e9
    print(f"texture Adam step R=2048: torch Adam + fill + repack {t_ref:.3f} ms, fused {t_new:.3f} ms ({gbs:.0f} GB/s algorithmic)")
    assert t_new < t_ref
