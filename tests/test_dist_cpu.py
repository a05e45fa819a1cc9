"""world_size-2 gloo tests (CPU) of the data-parallel host logic: view sharding and the flat
gradient bucket whose slices are the leaves' .grad, summed with one all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from texture_gs_b200.dist import GradBucket, shard_views


def test_shard_views_partitions_exactly():
    for n in (32, 7, 1, 0):
        for w in (1, 2, 3, 8):
            parts = [shard_views(n, w, r) for r in range(w)]
            flat = [v for p in parts for v in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_bucket_slices_are_the_grads_and_accumulate_in_place():
    a = torch.randn(5, 3, requires_grad=True)
    b = torch.randn(7, requires_grad=True)
    c = torch.randn(2, 2)                       # no grad: not in the bucket
    bk = GradBucket({"a": a, "b": b, "c": c, "none": None})
    assert set(bk.params) == {"a", "b"}
    ptr_a = a.grad.data_ptr()
    for _ in range(3):                            # three "views" accumulate into the same storage
        ((a * 2).sum() + (b * 3).sum()).backward()
    assert a.grad.data_ptr() == ptr_a == bk.flat.data_ptr()
    assert torch.allclose(bk.grads()["a"], torch.full((5, 3), 6.0))
    assert torch.allclose(bk.grads()["b"], torch.full((7,), 9.0))
    bk.zero()
    assert float(bk.flat.abs().sum()) == 0 and a.grad.data_ptr() == ptr_a


def test_fused_block_hands_out_the_bucket_storage_of_exact_leaves_only():
    """Inside ``bucket.fused()`` the rasterizer asks ``storage_for`` where to accumulate: the padded (6,R,R,4) storage for
    the texture, plain storage for the others, nothing for tensors that are not the bucket's leaves. One buffer for every
    stream (the kernels' accumulations are atomic), so there is nothing to fold before the all-reduce."""
    from texture_gs_b200.dist import current_fused_bucket
    tex = torch.randn(6, 4, 4, 3, requires_grad=True)
    x = torch.randn(9, 3, requires_grad=True)
    other = torch.randn(9, 3, requires_grad=True)
    bk = GradBucket({"texture": tex, "xyz": x})
    assert tex.grad.shape == tex.shape and current_fused_bucket() is None
    with bk.fused():
        assert current_fused_bucket() is bk
        buf, padded = bk.storage_for(tex)
        assert padded and buf.shape == (6, 4, 4, 4)
        buf[..., :3] += 2.0
        bx, px = bk.storage_for(x)
        assert not px
        bx += 10.0
        assert bk.storage_for(other) is None and bk.storage_for(x.detach()) is None       # only the exact leaf tensors
    assert current_fused_bucket() is None
    assert bk.storage_for(tex)[0].data_ptr() == bk.flat.data_ptr() + 4 * bk.offsets["texture"][0]
    assert bk.all_reduce() is None                                                              # no process group
    assert torch.allclose(tex.grad, torch.full_like(tex, 2.0)) and torch.allclose(x.grad, torch.full_like(x, 10.0))
    pad = bk.flat[bk.offsets["texture"][0]: bk.offsets["texture"][0] + bk.offsets["texture"][1]].view(-1, 4)[:, 3]
    assert float(pad.abs().max()) == 0.0
    bk.zero()
    assert float(bk.flat.abs().max()) == 0.0


def test_multi_stream_rendering_needs_a_bucket():
    from texture_gs_b200.dist import render_views_accumulate
    with pytest.raises(ValueError, match="GradBucket"):
        render_views_accumulate(None, None, [None], None, [0, 1, 2], None, bucket=None, streams=2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.randn(6, 4)                        # replicated "texture"
    x = torch.randn(3)                           # replicated "gaussian parameter"
    w.requires_grad_(True); x.requires_grad_(True)
    bk = GradBucket({"w": w, "x": x})
    views = shard_views(10, world, rank)
    for v in views:                              # a view's loss depends on its index
        ((w * (v + 1)).sum() + (x * x).sum() * (v + 1)).backward()
    bk.all_reduce()
    # expected: sum over all 10 views
    s = sum(v + 1 for v in range(10))
    ok = torch.allclose(bk.grads()["w"], torch.full((6, 4), float(s))) and \
        torch.allclose(bk.grads()["x"], 2 * x.detach() * s)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_single_process_sum():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


_QUIET_CHILD = r"""
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch, torch.distributed as dist
from texture_gs_b200.dist import init_process_group_quiet
rank = int(os.environ["RANK"])
print("before", flush=True)                                  # stdout, rank-tagged below
init_process_group_quiet("gloo", torch.device("cpu"))
os.write(1, b"")                                              # fd 1 is usable again
t = torch.ones(1) * (rank + 1)
dist.all_reduce(t)
print('{"rank": %d, "sum": %d}' % (rank, int(t.item())), flush=True)
dist.destroy_process_group()
"""


def test_quiet_init_restores_stdout_and_the_group_works(tmp_path):
    """bench.py's N>1 path: init + barrier with fd 1 parked on stderr, then exactly the JSON line on stdout."""
    import subprocess
    import sys
    from pathlib import Path
    port = _free_port()
    script = tmp_path / "child.py"
    script.write_text(_QUIET_CHILD)
    procs = []
    for r in range(2):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r), WORLD_SIZE="2",
                   REPO_ROOT=str(Path(__file__).resolve().parent.parent))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    for r, (p, (so, se)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, se
        assert so.splitlines() == ["before", '{"rank": %d, "sum": 3}' % r], (so, se)


def test_all_reduce_ranges_leave_the_excluded_parameters_alone():
    """bucket.all_reduce(exclude=("texture",)): the texture slice is reduced by DistTextureAdam's own kernel; everything
    else goes out as the contiguous ranges around it."""
    tex = torch.randn(6, 4, 4, 3, requires_grad=True)
    a = torch.randn(9, 3, requires_grad=True)
    b = torch.randn(5, requires_grad=True)
    bk = GradBucket({"a": a, "texture": tex, "b": b})
    r = bk.ranges_without(("texture",))
    o, n = bk.offsets["texture"]
    assert r == [(0, o), ((o + n + 63) // 64 * 64, bk.flat.numel())]
    assert bk.ranges_without(()) == [(0, bk.flat.numel())]
    covered = torch.zeros(bk.flat.numel(), dtype=torch.bool)
    for x, y in r:
        covered[x:y] = True
    for k in ("a", "b"):
        ko, kn = bk.offsets[k]
        assert bool(covered[ko:ko + kn].all())
    assert not bool(covered[o:o + n].any())


def test_dp_shard_partitions_the_texture_tiles():
    """texgs_dp_shard: contiguous tile ranges (1024 texels per tile) that cover every tile exactly once, sizes within one."""
    import ctypes as C
    from texture_gs_b200 import _lib as L
    lib = L.load()
    for n, world in ((6 * 2048 * 2048, 8), (6 * 96 * 96, 3), (1000, 4), (1024 * 7 + 5, 2)):
        tiles = (n + 1023) // 1024
        prev = 0
        for r in range(world):
            lo, hi = C.c_uint64(), C.c_uint64()
            assert lib.texgs_dp_shard(n, world, r, C.byref(lo), C.byref(hi)) == 0
            assert lo.value == prev and hi.value >= lo.value and hi.value - lo.value <= tiles // world + 1
            prev = hi.value
        assert prev == tiles
    assert lib.texgs_dp_shard(10, 2, 2, C.byref(lo), C.byref(hi)) != 0
