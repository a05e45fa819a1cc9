"""Reference-shaped training step on the headline config (stage 3 after iteration 10 000,
models/texture_gaussian3d.py:315-410): two renders per view (with SH, and active_sh_degree = 0), photometric
loss (1-l)*L1 + l*(1-SSIM) on both images, L1 on alpha, masked normal loss, bilateral normal smoothness
(the losses configs/texture_gaussian3d.yaml:77-88 enables), backward.
  A: as the reference does it — two rasterizer calls + the PyTorch loss formulation
  B: this repo's next-row pieces — one dual render (N2) + fused photometric and geometry losses (N3)
Prints one JSON line with ms per view for both."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import loss_ref as LR          # the reference's loss formulation (measurement harness only)
from texture_gs_b200 import uv_tex_render, uv_tex_render_dual
from texture_gs_b200.losses import photometric_loss, geometry_losses
from texture_gs_b200.scene import sphere_shell_scene, orbit_cameras

N, W, H, R = 500000, 1920, 1080, 2048
g = sphere_shell_scene(N, R, device="cuda")
cams = orbit_cameras(8, W, H, device="cuda")
bg = torch.zeros(3, device="cuda")
gen = torch.Generator().manual_seed(0)
gt = torch.rand(3, H, W, generator=gen).cuda()
gt_alpha = (torch.rand(1, H, W, generator=gen) > 0.2).float().cuda()
gt_norm = torch.nn.functional.normalize(torch.randn(3, H, W, generator=gen), dim=0).cuda()
lam, lam_nosh, lam_alpha, lam_norm, lam_nsm = 0.2, 2.0, 1.0, 0.1, 0.5


def step_reference_style(i):
    pkg = uv_tex_render(cams[i % 8], g, None, bg)
    la, ln, ls = LR.geometry_losses(pkg["alpha"], pkg["norm"], gt_alpha, gt_norm, gt)
    loss = LR.photometric_loss(pkg["render"], gt, lam)[0] + lam_alpha * la + lam_norm * ln + lam_nsm * ls
    deg = g.active_sh_degree
    g.active_sh_degree = 0
    img0 = uv_tex_render(cams[i % 8], g, None, bg)["render"]
    g.active_sh_degree = deg
    loss = loss + lam_nosh * LR.photometric_loss(img0, gt, lam)[0]
    loss.backward()
    g.zero_grad()


def step_fused(i):
    pkg = uv_tex_render_dual(cams[i % 8], g, None, bg)
    la, ln, ls = geometry_losses(pkg["alpha"], pkg["norm"], gt_alpha, gt_norm, gt)
    loss = photometric_loss(pkg["render"], gt, lam)[0] + lam_alpha * la + lam_norm * ln + lam_nsm * ls
    loss = loss + lam_nosh * photometric_loss(pkg["render_no_sh"], gt, lam)[0]
    loss.backward()
    g.zero_grad()


def timeit(fn, n=12, warm=4):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

a = timeit(step_reference_style)
b = timeit(step_fused)

# the same two steps followed by the texture's optimizer step (models/texture_gaussian3d.py:439-444)
from texture_gs_b200.dist import GradBucket
from texture_gs_b200.optim import TextureAdam
tex = g.get_texture
opt_ref = torch.optim.Adam([{"params": [tex], "lr": 0.0025}], lr=0.0, eps=1e-15)


def iter_reference_style(i):
    pkg = uv_tex_render(cams[i % 8], g, None, bg)
    la, ln, ls = LR.geometry_losses(pkg["alpha"], pkg["norm"], gt_alpha, gt_norm, gt)
    loss = LR.photometric_loss(pkg["render"], gt, lam)[0] + lam_alpha * la + lam_norm * ln + lam_nsm * ls
    deg = g.active_sh_degree
    g.active_sh_degree = 0
    img0 = uv_tex_render(cams[i % 8], g, None, bg)["render"]
    g.active_sh_degree = deg
    (loss + lam_nosh * LR.photometric_loss(img0, gt, lam)[0]).backward()
    opt_ref.step()
    g.zero_grad()                                  # zero_grad(set_to_none=True)


c = timeit(iter_reference_style)
g.zero_grad()
bucket = GradBucket({"texture": tex})
opt_new = TextureAdam([{"params": [tex], "lr": 0.0025}], lr=0.0, eps=1e-15, zero_grad_in_step=True)


def iter_fused(i):
    with bucket.fused():
        pkg = uv_tex_render_dual(cams[i % 8], g, None, bg)
        la, ln, ls = geometry_losses(pkg["alpha"], pkg["norm"], gt_alpha, gt_norm, gt)
        loss = photometric_loss(pkg["render"], gt, lam)[0] + lam_alpha * la + lam_norm * ln + lam_nsm * ls
        (loss + lam_nosh * photometric_loss(pkg["render_no_sh"], gt, lam)[0]).backward()
    opt_new.step()
    for k, v in g.tensors().items():
        if v is not None and v is not tex:
            v.grad = None


d = timeit(iter_fused)
print(json.dumps({"config": "500k / 1080p / R2048, stage-3 step (2 images, L1+SSIM on both, alpha L1, normal loss, normal smoothness, backward)",
                  "two_renders_torch_losses_ms": round(a, 3), "dual_render_fused_losses_ms": round(b, 3), "speedup": round(a / b, 2),
                  "with_texture_adam": {"torch_adam_ms": round(c, 3), "fused_bucket_texture_adam_ms": round(d, 3), "speedup": round(c / d, 2)}}))
