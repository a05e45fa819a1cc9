"""Data-parallel texture step on >= 2 GPUs:  torchrun --nproc-per-node N tests/gpu_dist_adam.py [R] [views_per_rank]
Every rank renders its own views of a small scene into a symmetric gradient bucket; DistTextureAdam (one fused kernel per
rank: peer / multicast gradient pull + Adam on the owned shard + parameter push) must leave on EVERY rank the texture that
the reference recipe leaves — NCCL all-reduce of the gradient followed by the single-GPU TextureAdam step — and the same
Adam moments on the owned shard. Both the multicast (NVLS) and the peer-to-peer path are run when the fabric offers
multicast. Prints one JSON line on rank 0; exit code 1 on mismatch."""
import json, os, sys
from pathlib import Path
import torch
import torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from texture_gs_b200 import invalidate_packed_cache, uv_tex_render
from texture_gs_b200.dist import DistTextureAdam, GradBucket, init_process_group_quiet, render_views_accumulate
from texture_gs_b200.optim import TextureAdam
from texture_gs_b200.scene import orbit_cameras, output_cotangents, sphere_shell_scene

R = int(sys.argv[1]) if len(sys.argv) > 1 else 96
VPR = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
init_process_group_quiet("nccl", dev)
W, H, N = 320, 192, 20000
cams = orbit_cameras(world * VPR, W, H, seed=1, device=dev)
cot = output_cotangents(H, W, seed=3, device=dev)
bg = torch.zeros(3, device=dev)
views = list(range(rank * VPR, (rank + 1) * VPR))
report = {"world": world, "R": R, "paths": {}}
ok = True
for path in ("multicast", "peer"):
    g = sphere_shell_scene(N, R, sh_degree=3, seed=0, device=dev)
    g_ref = sphere_shell_scene(N, R, sh_degree=3, seed=0, device=dev)
    bucket = GradBucket(g.tensors(), symmetric_group=dist.group.WORLD)
    opt = DistTextureAdam(g.get_texture, bucket, lr=0.0025, eps=1e-15, use_multicast=(path == "multicast"))
    if path == "multicast" and not opt.multicast:
        report["paths"][path] = "no multicast mapping on this fabric"
        continue
    bucket_ref = GradBucket(g_ref.tensors())
    opt_ref = TextureAdam([g_ref.get_texture], lr=0.0025, eps=1e-15)
    worst = 0.0
    for step in range(3):
        for b, gg in ((bucket, g), (bucket_ref, g_ref)):
            invalidate_packed_cache()
            b.zero()
            render_views_accumulate(uv_tex_render, gg, cams, cot, views, bg, bucket=b, streams=2)
        works = bucket.all_reduce(exclude=("texture",), async_op=True)
        opt.step()
        for w in works or []:
            w.wait()
        bucket_ref.all_reduce()
        opt_ref.step()
        torch.cuda.synchronize()
        d_tex = float((g.get_texture.detach() - g_ref.get_texture.detach()).abs().max())
        lo, m, v = opt.state_shard()
        n_own = min(opt.tile_hi * 1024, opt.n_texels) - lo
        st = opt_ref.state[g_ref.get_texture]
        d_m = float((m[:n_own * 3] - st["exp_avg"].reshape(-1)[lo * 3:(lo + n_own) * 3]).abs().max()) if n_own > 0 else 0.0
        d_v = float((v[:n_own * 3] - st["exp_avg_sq"].reshape(-1)[lo * 3:(lo + n_own) * 3]).abs().max()) if n_own > 0 else 0.0
        d_rest = max(float((bucket.grads()[k] - bucket_ref.grads()[k]).abs().max()) for k in bucket.params if k != "texture")
        moved = float((g.get_texture.detach() - sphere_shell_scene(8, R, seed=0, device=dev).get_texture).abs().max()) if step == 0 else 1.0
        worst = max(worst, d_tex)
        t = torch.tensor([d_tex, d_m, d_v, d_rest, -moved], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d_tex, d_m, d_v, d_rest, moved = t.tolist()
        # atomics order differs between the two renders of a rank: gradients agree to fp32 summation noise, Adam's
        # m / sqrt(v) turns that into at most ~lr * 1e-3 on a texel
        good = d_tex <= 2.5e-5 and d_m <= 1e-3 and -moved > 1e-4
        ok = ok and good
        report["paths"].setdefault(path, []).append({"step": step, "max|tex - ref|": d_tex, "max|m - ref|": d_m, "max|v - ref|": d_v,
                                                       "max|other grads - ref|": d_rest, "ok": good})
    del opt, bucket
if rank == 0:
    report["ok"] = ok
    print(json.dumps(report))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
