#!/usr/bin/env python
"""bench.py — views/s forward+backward of the Texture-GS rasterizer hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the C + OpenMP oracle (port of the
                                                             # reference algorithm) on all host cores, whole views

Workload (N=1 and N>1 alike): BASELINE.json configs[2]/[3] — 500k synthetic Gaussians ("sphere-shell",
seed 0), 1920x1080, cube texture 6x2048^2x3, sh_degree 3; a *step* is one batch of 32 views
(forward + backward with dense cotangents on all four outputs, gradients accumulated into one flat
bucket). With N GPUs the 32 views are sharded over the ranks and the bucket is summed with one NCCL
all-reduce per step (strong scaling: total work fixed). value = 32*K / time, time = max over ranks
of CUDA-event time around the K steps bracketed by barrier + synchronize.

The JSON line also carries: e2e (same step driven from HOST buffers: per-view cotangent images are
copied from pinned host memory, the per-step loss is read back), roofline (dominant kernel,
algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json), cpu_baseline (the oracle timed on
a bounded sample), clocks (nvidia-smi sampled during the timed region), gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

VIEWS_PER_STEP = 32
KERNELS_PER_VIEW = 8   # (+1 texgs_pack_texture_kernel per step) preprocess_fwd, scan_tiles, scatter_pairs, sort_tiles_small,
                       # sort_tiles, render_fwd, render_bwd, preprocess_bwd


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_500k_1080p")
    ap.add_argument("--views", type=int, default=VIEWS_PER_STEP)
    ap.add_argument("--streams", type=int, default=1,
                    help="experimental: CUDA streams per rank that render alternate views concurrently (one gradient-bucket replica each)")
    ap.add_argument("--fwd-ilp2", action="store_true", help="experimental: forward blend kernel with two splats per half-warp per iteration")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-tiles", type=int, default=0, help="ignored (kept for old command lines): the CPU arm renders whole views")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, clk_max = float(f[0]), float(f[1])
            except ValueError:
                continue
            sm.append(clk); mx.append(clk_max)
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass                      # power.draw can read [N/A]; the clocks and the throttle reasons still count
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(N, M, V, K, U, P, R, views_per_step=VIEWS_PER_STEP):
    """SURVEY.md §8d / DESIGN.md §4: algorithmic HBM bytes per view, split per kernel.
    N Gaussians, M SH-rest coeffs, V visible, K (tile,Gaussian) pairs, U unique texels touched,
    P pixels, R face resolution. Records are 128 B, accumulators 20 floats."""
    b = {}
    b["preprocess_fwd"] = N * (92 + 12 * M) + V * (128 + 8 + 4) + K * 4
    b["scan_tiles"] = 0
    b["scatter_pairs"] = V * 12 + K * 8
    b["sort_tiles"] = K * 8 + K * 12
    b["render_fwd"] = K * (4 + 128) + U * 12 + P * 40
    b["render_bwd"] = P * 40 + K * (4 + 128) + U * 12 + 2 * U * 12 + 2 * V * 80
    # per view: the accumulator clear; the dense texture-gradient zero fill happens once per step (GradBucket.zero()),
    # outside the per-view kernels, and is charged to the view at 1 / views_per_step
    b["bwd_clear"] = N * 96 + 6 * R * R * 16 // max(1, views_per_step)
    b["preprocess_bwd"] = N * (92 + 12 * M) + V * 80 + N * (68 + 12 * M)
    return b


def roofline_report(wl, views_per_step, world, value, stage_ms, stats, U, clock_rec, sms):
    """The ``roofline`` object of the JSON line (pure host arithmetic, unit-tested on CPU): per-kernel algorithmic bytes
    over the CUDA-event durations measured in the timed region, the dominant kernel against the measured HBM peak, its
    DRAM traffic and instruction count from the committed ncu captures."""
    hbm_peak, peak_src = measured_peaks()
    N, M = wl.n_gaussians, 15
    ab = algorithmic_bytes(N, M, stats.num_visible, stats.num_pairs, U, wl.width * wl.height, wl.tex_res, views_per_step)
    per_kernel = {}
    for k, b in ab.items():
        if k in stage_ms and stage_ms[k] > 0:
            gbs = b / (stage_ms[k] * 1e-3) / 1e9
            per_kernel[k] = {"ms": round(stage_ms[k], 4), "alg_mb": round(b / 1e6, 2), "gbs": round(gbs, 1), "frac": round(gbs / hbm_peak, 4)}
    whole = {"alg_mb_per_view": round(sum(ab.values()) / 1e6, 1),
             "gbs": round(sum(ab.values()) * value / 1e9 / max(world, 1), 1),
             "frac": round(sum(ab.values()) * value / 1e9 / max(world, 1) / hbm_peak, 4)}
    counts = {"V": stats.num_visible, "K": stats.num_pairs, "U": U, "max_tile_len": stats.max_tile_len}
    if not per_kernel:          # no stage events recorded (profiling slots exhausted): report the whole path only
        return {"bound": "hbm", "kernel": None, "achieved": whole["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": whole["frac"],
                "traffic": None, "peak_source": peak_src, "per_kernel": {}, "whole_path": whole, "counts": counts}
    dom = max((k for k in per_kernel), key=lambda k: per_kernel[k]["ms"])
    traffic = None
    tfile = ROOT / "profiles" / "ncu_traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(dom)
        except Exception:
            traffic = None
    # second roof for the dominant kernel: instruction issue. Static warp-instruction count of the ncu capture of the same
    # kernels and config (profiles/ncu_issue.json) over the duration measured live; peak = SMs x 4 schedulers x SM clock
    issue = None
    try:
        ifile = ROOT / "profiles" / "ncu_issue.json"
        if ifile.exists() and clock_rec and clock_rec.get("sm_mhz"):
            rec = json.loads(ifile.read_text()).get(dom)
            if rec:
                rate = rec["warp_instructions"] / (per_kernel[dom]["ms"] * 1e-3) / 1e9
                peak = sms * 4 * clock_rec["sm_mhz"] * 1e6 / 1e9
                issue = {"warp_instructions_per_launch": rec["warp_instructions"], "achieved_ginst_s": round(rate, 1),
                         "peak_ginst_s": round(peak, 1), "frac": round(rate / peak, 4),
                         "active_threads_per_instruction": rec.get("active_threads_per_instruction"),
                         "source": "instruction count from the ncu capture in profiles/ncu_issue.json, duration measured live"}
    except Exception:
        issue = None
    return {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
            "frac": per_kernel[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ab[dom], "issue": issue, "per_kernel": per_kernel,
            "whole_path": whole, "counts": counts}


def make_scene(wl, device, requires_grad=True):
    from texture_gs_b200.scene import orbit_cameras, output_cotangents, sphere_shell_scene
    g = sphere_shell_scene(wl.n_gaussians, wl.tex_res, sh_degree=3, seed=0, device=device, requires_grad=requires_grad)
    cams = orbit_cameras(VIEWS_PER_STEP, wl.width, wl.height, seed=1, device=device)
    cot = output_cotangents(wl.height, wl.width, seed=3, device=device)
    return g, cams, cot


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle on host cores (bounded sample)
# ---------------------------------------------------------------------------------------------

_cpu_cache = {}


def cpu_oracle_views_per_s(wl, sample_tiles: int = 0, backward: bool = True):
    """CPU arm: the C + OpenMP oracle (oracle/raster_c.c, float32 build) renders ONE WHOLE view of the workload —
    all Gaussians, all tiles, forward + backward — on all host threads OpenMP gives it. No sampling, no extrapolation
    (the torch oracle needed both; at 40-170 s per view it could only be timed on 5 % of the tiles).
    Returns (views/s, description, threads, seconds)."""
    from oracle import raster_c
    from oracle.raster_ref import RasterSettings
    from texture_gs_b200.scene import orbit_cameras, output_cotangents, sphere_shell_scene
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    threads = int(os.environ.get("OMP_NUM_THREADS", 0)) or avail
    key = (wl.name, backward)
    if key not in _cpu_cache:
        g = sphere_shell_scene(wl.n_gaussians, wl.tex_res, sh_degree=3, seed=0, device="cpu", requires_grad=False)
        cams = orbit_cameras(VIEWS_PER_STEP, wl.width, wl.height, seed=1)
        cot = output_cotangents(wl.height, wl.width, seed=3)
        raster_c.build(torch.float32)
        _cpu_cache[key] = (g.tensors(), cams, cot, [0])
    t, cams, cot, counter = _cpu_cache[key]
    cam = cams[counter[0] % len(cams)]              # a different camera every call, like the GPU arm
    counter[0] += 1
    st = RasterSettings(wl.height, wl.width, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), torch.zeros(3), 1.0,
                        cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center)
    t0 = time.perf_counter()
    out = raster_c.rasterize(t["xyz"], t["shs"], t["opacity"], t["scaling"], t["rotation"], t["uvs"], t["grad_uvs"], t["texture"], st,
                             cotangents=cot if backward else None, dtype=torch.float32, threads=threads)
    dt = time.perf_counter() - t0
    aux = out[-1]
    desc = (f"C + OpenMP oracle (oracle/raster_c.c, float32) fwd{'+bwd' if backward else ''} of 1 whole view: {wl.n_gaussians} Gaussians, "
            f"{wl.width}x{wl.height}, {aux['num_pairs']} (tile,Gaussian) pairs, {aux['num_blend']} blended contributions in {dt:.2f} s "
            f"on {threads} threads (no sampling, no extrapolation)")
    return 1.0 / dt, desc, threads, dt


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    desc = ""
    threads = 1
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, desc, threads, dt = cpu_oracle_views_per_s(wl)
        if i >= args.warmup:
            vals.append(v)
        if time.perf_counter() - t_all > 240:
            break
    if not vals:
        vals = [v]
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": "views/s fwd+bwd", "value": value, "unit": "views/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1000.0 * VIEWS_PER_STEP / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "gaussians": wl.n_gaussians, "width": wl.width, "height": wl.height,
                       "tex_res": wl.tex_res, "views_per_step": VIEWS_PER_STEP, "sh_degree": 3,
                       "parallelism": f"host CPU, {threads} OpenMP threads (rank 0 only)",
                       "l2_policy": "n/a (CPU arm)"},
            "cpu_baseline": {"value": value, "unit": "views/s", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

def main():
    args = parse()
    from texture_gs_b200.scene import WORKLOADS
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch.distributed as dist
    from texture_gs_b200 import _lib, last_stats, uv_tex_render
    from texture_gs_b200.dist import GradBucket, render_views_accumulate, shard_views
    from texture_gs_b200.profiling import StageTimer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from texture_gs_b200.dist import init_process_group_quiet
        init_process_group_quiet("nccl", dev)        # keeps NCCL's version banner off stdout (one JSON line only)
    _lib.load()
    if args.fwd_ilp2:
        from texture_gs_b200 import rasterizer as _rz
        _rz.FWD_ILP2 = True

    g, cams, cot = make_scene(wl, dev)
    bg = torch.zeros(3, device=dev)
    params = {k: v for k, v in g.tensors().items()}
    bucket = GradBucket(params, replicas=max(1, args.streams))
    views = shard_views(args.views, world, rank)
    timer = StageTimer(capacity=max(1, len(views)) * max(1, args.steps), device=dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from texture_gs_b200 import invalidate_packed_cache

    def step(tm=None):
        # one texture update per step: the packed (6,R,R,4) copy is rebuilt once per 32-view batch
        invalidate_packed_cache()
        bucket.zero()
        render_views_accumulate(uv_tex_render, g, cams, cot, views, bg, timer=tm, bucket=bucket, streams=args.streams)
        bucket.all_reduce()

    for _ in range(args.warmup):
        step()
    sync_all()
    stats = last_stats()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step(timer)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clock_rec = clocks.stop() if rank == 0 else None
    stage_ms = timer.summary()
    value = args.views * args.steps / (ms / 1e3)

    # ---- e2e: same step, driven from host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, wl, g, cams, cot, bg, bucket, views, dev, world, sync_all)
        except Exception as e:              # keep the device-resident measurement; the line then says why e2e is missing
            if world > 1:
                raise                       # a rank that drops out of the collectives would hang the others
            e2e = {"value": None, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(e)[:300]}

    # ---- unique texels touched by one view (for the algorithmic byte count) ---------------------
    bucket.zero()
    render_views_accumulate(uv_tex_render, g, cams, [torch.ones_like(c) for c in cot], views[:1] or [0], bg, bucket=bucket)
    U = int((bucket.grads()["texture"].abs().sum(dim=-1) > 0).sum().item())
    bucket.zero()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    roofline = roofline_report(wl, args.views, world, value, stage_ms, stats, U, clock_rec, sms)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            vs = [cpu_oracle_views_per_s(wl) for _ in range(3)]       # a few seconds each; first call builds the scene
            v, desc, threads, _ = max(vs, key=lambda r: r[0])
            cpu = {"value": v, "unit": "views/s", "cores": threads, "kind": "port", "sample": desc + "; best of 3 views"}
        except Exception as e:                                        # never lose the GPU measurement to the CPU leg
            cpu = {"value": None, "unit": "views/s", "cores": 0, "kind": "port", "sample": ("unavailable: " + repr(e))[:300]}

    line = {"metric": "views/s fwd+bwd", "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "gaussians": wl.n_gaussians, "width": wl.width, "height": wl.height,
                       "tex_res": wl.tex_res, "views_per_step": args.views, "sh_degree": 3, "streams_per_rank": max(1, args.streams),
                       "forward_kernel": "ilp2 (experimental)" if args.fwd_ilp2 else "default",
                       "parallelism": f"dp{world} (views sharded, 1 all-reduce of {bucket.nbytes / 1e6:.0f} MB/step)" if world > 1 else "single GPU",
                       "l2_policy": "inputs larger than L2 (texture 302 MB + records 64 MB, a different camera every view)"},
            "clocks": clock_rec, "e2e": e2e, "gpu_launches": (KERNELS_PER_VIEW * len(views) + 1) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, wl, g, cams, cot_dev, bg, bucket, views, dev, world, sync_all):
    """The same step through the public operator with HOST inputs: for every view the four dense
    cotangent images (the stand-in for the per-view supervision images of the reference's training
    loop, train.py:147-149) are copied from pinned host memory on a side stream (double buffered),
    the per-step scalar result is read back to the host. All copies are inside the timed region."""
    import torch.distributed as dist
    from texture_gs_b200 import uv_tex_render
    host = [c.cpu().pin_memory() for c in cot_dev]
    bytes_view = sum(h.numel() * 4 for h in host) + (16 + 16 + 3) * 4
    copy_stream = torch.cuda.Stream(dev)
    bufs = [[torch.empty_like(c) for c in cot_dev] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    result_host = torch.zeros(1).pin_memory()

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[slot])
            for d, h in zip(bufs[slot], host):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    from texture_gs_b200 import invalidate_packed_cache

    primed = [False]      # view 0 of the NEXT step is uploaded while the last view of this step renders (loader prefetch)

    def step():
        invalidate_packed_cache()
        bucket.zero()
        total = torch.zeros((), device=dev)
        main = torch.cuda.current_stream(dev)
        if views and not primed[0]:
            for s in range(2):
                free[s].record(main)
            upload(0)
        nv = len(views)
        for i, v in enumerate(views):
            slot = i & 1
            if i + 1 < nv:
                upload((i + 1) & 1)
            elif nv % 2 == 0:
                upload(0)                      # next step's first view (slot 0 was released after view nv-2)
            main.wait_event(ready[slot])
            cam = cams[v % len(cams)]
            with bucket.fused():
                pkg = uv_tex_render(cam, g, None, bg)
                outs = [pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]]
                with torch.no_grad():
                    total += sum(torch.dot(o.detach().reshape(-1), c.reshape(-1)) for o, c in zip(outs, bufs[slot]))
                torch.autograd.backward(outs, list(bufs[slot]))
            free[slot].record(main)
        primed[0] = nv > 0 and nv % 2 == 0
        bucket.all_reduce()
        result_host.copy_(total.reshape(1), non_blocking=True)

    for _ in range(max(3, args.warmup)):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3 * 0.0)
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    return {"value": args.views * args.steps / (ms / 1e3), "unit": "views/s", "h2d_bytes_per_step": bytes_view * len(views),
            "d2h_bytes_per_step": 4, "ms_per_step": ms / args.steps, "result": float(result_host.item())}


if __name__ == "__main__":
    main()
