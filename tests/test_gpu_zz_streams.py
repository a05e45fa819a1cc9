"""Multi-stream view pipeline (texture_gs_b200.dist.render_views_accumulate(..., streams=n)): same
gradients as the single-stream loop. Runs last in the GPU suite (file name)."""
import pytest
import torch

from util import rel_err
from texture_gs_b200.scene import orbit_cameras, output_cotangents, sphere_shell_scene

pytestmark = pytest.mark.gpu


def test_two_stream_accumulation_equals_single_stream():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected (-m gpu) but no CUDA device is visible")
    from texture_gs_b200 import invalidate_packed_cache, uv_tex_render
    from texture_gs_b200.dist import GradBucket, render_views_accumulate
    cams = orbit_cameras(5, 160, 96, seed=11, device="cuda")
    cot = output_cotangents(96, 160, seed=12, device="cuda")
    bg = torch.tensor([0.1, 0.0, 0.2], device="cuda")
    res = []
    for streams in (1, 2, 3):
        g = sphere_shell_scene(3000, 32, sh_degree=3, seed=10, device="cuda")
        bk = GradBucket(g.tensors())
        for _ in range(2):                           # second pass: caches are warm
            invalidate_packed_cache()
            bk.zero()
            render_views_accumulate(uv_tex_render, g, cams, cot, range(5), bg, bucket=bk, streams=streams)
            bk.all_reduce()
        torch.cuda.synchronize()
        res.append({k: v.detach().clone() for k, v in bk.grads().items()})
    for other in res[1:]:
        for k in res[0]:
            assert rel_err(other[k], res[0][k]) < 1e-5, k
