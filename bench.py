#!/usr/bin/env python
"""bench.py — views/s of the Texture-GS rasterizer hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the C + OpenMP oracle (port of the
                                                             # reference algorithm) on all host cores, whole views
    python bench.py --workload cfg1_300k_800x600             # forward-only lines (BASELINE configs[1], [4])

Default workload (N=1 and N>1 alike): BASELINE.json configs[2]/[3] — 500k synthetic Gaussians ("sphere-shell",
seed 0), 1920x1080, cube texture 6x2048^2x3, sh_degree 3; a *step* is one batch of 32 views (forward + backward with
dense cotangents on all four outputs, gradients accumulated into one flat bucket), rendered on ``--streams`` CUDA
streams per rank. With N GPUs the 32 views are sharded over the ranks and the bucket is summed with one NCCL
all-reduce per step (strong scaling: total work fixed). value = 32*K / time, time = max over ranks of CUDA-event time
around the K steps bracketed by barrier + synchronize.

The JSON line also carries: e2e (the training-shaped step driven from HOST buffers: per view the uint8 ground-truth
image, uint8 alpha mask and int8 normal map are copied from pinned host memory, the losses of the reference's
compute_loss are evaluated by the fused loss kernels, the per-step loss is read back), roofline (dominant kernel,
algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json), cpu_baseline (the oracle timed on a bounded
sample), clocks (nvidia-smi sampled during the timed region), gpu_launches, grad_checksum.
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


DEFAULT_STREAMS = 4    # views in flight per rank (profiles/r2_variants.md: 1 -> 371, 2 -> 396, 3 -> 405 views/s; 4 = 3 at one GPU and
                       # divides the 4 views a rank renders at 8 GPUs: no straggler view running alone)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_500k_1080p")
    ap.add_argument("--views", type=int, default=VIEWS_PER_STEP)
    ap.add_argument("--streams", type=int, default=DEFAULT_STREAMS,
                    help="CUDA streams per rank that render alternate views concurrently (they share the gradient bucket: atomic adds)")
    ap.add_argument("--no-optimizer", action="store_true",
                    help="leave the texture's Adam step out of the step and (N > 1) all-reduce the whole bucket with NCCL instead of "
                         "the fused reduce + Adam + broadcast kernel (the round-1 definition of the step)")
    ap.add_argument("--no-multicast", action="store_true", help="N > 1: peer loads / stores instead of NVSwitch multimem in the fused texture step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage-pass", action="store_true", help="skip the single-stream per-kernel timing pass (roofline.per_kernel)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arm (default: every CPU in the affinity mask)")
    return ap.parse_args()


def workload_config(wl, views):
    """The ``config`` object — the SAME keys and values in both arms (what is being measured, nothing about how)."""
    return {"workload": wl.name, "gaussians": wl.n_gaussians, "width": wl.width, "height": wl.height, "tex_res": wl.tex_res,
            "views_per_step": views, "sh_degree": 3, "pass": "fwd+bwd" if wl.backward else "fwd",
            "renders_per_view": wl.renders_per_view}


def metric_name(wl):
    return "views/s fwd+bwd" if wl.backward else "views/s fwd"


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


def algorithmic_bytes(N, M, V, K, U, P, R, views_per_step=VIEWS_PER_STEP, backward=True):
    """algorithmic HBM bytes per view, split per kernel, with the survey's per-unit sizes (NOT the
    sizes of this implementation's buffers: the 128-byte record, the 16-byte texel and the 24-float accumulator row
    move more than this; that excess shows up as ``traffic`` > algorithmic).
    N Gaussians, M SH-rest coeffs, V visible, K (tile,Gaussian) pairs, U unique texels touched, P pixels, R face res.
      forward  = N(92+12M) + 64V + 24K + 116K + 12U + 40P
      backward = 40P + 116K + 12U + 72R^2 [zero fill, once per STEP] + 24U + 136V + N(92+12M) + N(68+12M)"""
    b = {}
    b["preprocess_fwd"] = N * (92 + 12 * M) + V * 64
    b["scan_tiles"] = 0
    b["scatter_pairs"] = K * 12                      # key + value written once ...
    b["sort_tiles"] = K * 12                         # ... and read once
    b["render_fwd"] = K * 116 + U * 12 + P * 40
    if backward:
        b["render_bwd"] = P * 40 + K * 116 + U * 12 + 2 * U * 12 + 2 * V * 68
        b["bwd_clear"] = 6 * R * R * 12 // max(1, views_per_step)
        b["preprocess_bwd"] = N * (92 + 12 * M) + N * (68 + 12 * M)
    return b


def roofline_report(wl, views_per_step, world, value, stage_ms, stats, U, clock_rec, sms, timing_note=None):
    """The ``roofline`` object of the JSON line (pure host arithmetic, unit-tested on CPU): per-kernel algorithmic bytes
    over the CUDA-event durations measured in the timed region, the dominant kernel against the measured HBM peak, its
    DRAM traffic and instruction count from the committed ncu captures."""
    hbm_peak, peak_src = measured_peaks()
    N, M = wl.n_gaussians, 15
    ab = algorithmic_bytes(N, M, stats.num_visible, stats.num_pairs, U, wl.width * wl.height, wl.tex_res, views_per_step, wl.backward)
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
            "whole_path": whole, "counts": counts, "timing": timing_note}


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


def cpu_threads(requested: int = 0) -> int:
    """Threads of the CPU arm: every CPU of the affinity mask. ``OMP_NUM_THREADS`` is deliberately NOT consulted —
    torch.distributed.run injects OMP_NUM_THREADS=1 into every rank, which made the round-1 CPU arm single-threaded
    under torchrun; ``--cpu-threads`` / TEXGS_CPU_THREADS override."""
    if requested > 0:
        return requested
    env = int(os.environ.get("TEXGS_CPU_THREADS", "0") or 0)
    if env > 0:
        return env
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle_views_per_s(wl, threads: int = 0):
    """CPU arm: the C + OpenMP oracle (oracle/raster_c.c, float32 build) renders ONE WHOLE view of the workload —
    all Gaussians, all tiles, forward (+ backward where the workload has one) — on ``threads`` host threads.
    No tile sampling, no extrapolation. Returns (views/s, description, threads, seconds)."""
    from oracle import raster_c
    from oracle.raster_ref import RasterSettings
    from texture_gs_b200.scene import orbit_cameras, output_cotangents, sphere_shell_scene
    threads = cpu_threads(threads)
    backward = wl.backward
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
    reps = wl.renders_per_view if not backward else 1          # retexture.py renders every view twice (with SH, degree 0)
    for _ in range(reps):
        out = raster_c.rasterize(t["xyz"], t["shs"], t["opacity"], t["scaling"], t["rotation"], t["uvs"], t["grad_uvs"], t["texture"], st,
                                 cotangents=cot if backward else None, dtype=torch.float32, threads=threads)
    dt = time.perf_counter() - t0
    aux = out[-1]
    desc = (f"C + OpenMP oracle (oracle/raster_c.c, float32) fwd{'+bwd' if backward else ''} of 1 whole view"
            f"{' (x%d renders)' % reps if reps > 1 else ''}: {wl.n_gaussians} Gaussians, "
            f"{wl.width}x{wl.height}, {aux['num_pairs']} (tile,Gaussian) pairs, {aux['num_blend']} blended contributions in {dt:.2f} s "
            f"on {threads} threads (no tile sampling, no extrapolation)")
    return 1.0 / dt, desc, threads, dt


def run_reference(args, wl):
    """``--impl reference``: rank 0 only. A timed *step* of this arm is a bounded sample of the workload's 32-view step:
    ONE whole view (a different camera every step), so that K steps + W warm-ups end within minutes on host cores;
    ``ms_per_step`` is the measured mean of those K sampled steps (not derived), every requested step is run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads(args.cpu_threads)
    secs, desc = [], ""
    for i in range(args.warmup + args.steps):
        _v, desc, threads, dt = cpu_oracle_views_per_s(wl, threads)
        if i >= args.warmup:
            secs.append(dt)
    mean_s = sum(secs) / max(1, len(secs))
    value = 1.0 / mean_s if secs else 0.0
    line = {"impl": "reference", "metric": metric_name(wl), "value": value, "unit": "views/s", "n_gpus": args.gpus,
            "steps": len(secs), "warmup": args.warmup, "ms_per_step": 1000.0 * mean_s,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, args.views),
            "impl_notes": {"parallelism": f"host CPU, {threads} OpenMP threads (rank 0 only; OMP_NUM_THREADS ignored, see cpu_threads())",
                           "step_sample": "each timed step = 1 whole view of the 32-view step (bounded sample); ms_per_step is per sampled step",
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
    from texture_gs_b200 import _lib, last_stats, uv_tex_render, uv_tex_render_dual
    from texture_gs_b200.dist import GradBucket, bind_to_gpu_numa_node, render_views_accumulate, shard_views
    from texture_gs_b200.profiling import StageTimer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:
        full_affinity = os.sched_getaffinity(0)
    except AttributeError:
        full_affinity = None
    numa_node = bind_to_gpu_numa_node(dev)          # before any pinned allocation: host buffers land next to the GPU
    if world > 1:
        from texture_gs_b200.dist import init_process_group_quiet
        init_process_group_quiet("nccl", dev)        # keeps NCCL's version banner off stdout (one JSON line only)
    _lib.load()

    bwd = wl.backward
    streams = max(1, args.streams)
    g, cams, cot = make_scene(wl, dev, requires_grad=bwd)
    bg = torch.zeros(3, device=dev)
    # retexture.py renders every view twice (with SH, then active_sh_degree = 0): one dual render here
    render_fn = uv_tex_render_dual if (not bwd and wl.renders_per_view == 2) else uv_tex_render
    use_opt = bwd and not args.no_optimizer
    views = shard_views(args.views, world, rank)
    # the texture's optimizer (models/texture_gaussian3d.py:139-143: Adam, lr = tex_lr 0.0025, eps 1e-15) is part of the step:
    # one GPU runs the fused TextureAdam kernel; N GPUs run ONE kernel each that pulls + adds the ranks' partial texture gradients
    # over NVLink (multimem through the NVSwitch when available), updates the owned 1/N of the texels and pushes them to every rank —
    # instead of all-reducing 403 MB and repeating the same update N times. The per-Gaussian gradients (118 MB) go through NCCL.
    params = {k: v for k, v in g.tensors().items()}
    bucket, opt, fused_note = None, None, None
    if use_opt and world > 1:
        ok = 1
        try:
            if os.environ.get("TEXGS_BENCH_NO_SYMM"):       # test hook: exercise the fallback on a box that has symmetric memory
                raise RuntimeError("TEXGS_BENCH_NO_SYMM is set")
            from texture_gs_b200.dist import DistTextureAdam
            bucket = GradBucket(params, symmetric_group=dist.group.WORLD)
            opt = DistTextureAdam(g.get_texture, bucket, lr=0.0025, eps=1e-15, use_multicast=False if args.no_multicast else None)
        except Exception as e:      # no symmetric memory / peer access on this box: every rank all-reduces and steps on its own
            ok, fused_note = 0, "fused multi-GPU texture step unavailable (%s): NCCL all-reduce + TextureAdam on every rank" % repr(e)[:200]
        agree = torch.tensor([ok], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) == 0:
            bucket, opt = None, None
            fused_note = fused_note or "fused multi-GPU texture step unavailable on another rank: NCCL all-reduce + TextureAdam on every rank"
    if bwd and bucket is None:
        bucket = GradBucket(params)
    if use_opt and opt is None:
        from texture_gs_b200.optim import TextureAdam
        opt = TextureAdam([g.get_texture], lr=0.0025, eps=1e-15)
    fused_dp = opt is not None and hasattr(opt, "state_shard")
    tex0_l1 = float(g.get_texture.detach().abs().sum().item())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from texture_gs_b200 import invalidate_packed_cache

    from texture_gs_b200.rasterizer import ensure_packed_texture

    def finish_step():
        """What follows the last view of a step: reduce the gradients over the ranks and (unless --no-optimizer) update the texture."""
        if not bwd:
            return
        if opt is None:
            bucket.all_reduce()
        elif world > 1 and not fused_dp:
            bucket.all_reduce()                        # fallback: plain NCCL all-reduce, the same Adam step on every rank
            opt.step()
        elif world > 1:
            works = bucket.all_reduce(exclude=("texture",), async_op=True)      # per-Gaussian gradients: NCCL, concurrently
            opt.step()
            ensure_packed_texture(g.get_texture)       # repack of the pushed texels: under the tail of the NCCL collective
            for w in works or []:
                w.wait()
        else:
            opt.step()

    fin_ev = []            # (start, end) CUDA events around finish_step() of the timed steps
    # The tail of a batch (gradient reduction, optimizer, repack of the new texels, clearing the bucket) runs on the main
    # stream; the view streams of the NEXT batch fork from the moment this batch's views had joined (``joined``) and only
    # their render kernels wait for the tail (``tail_done``): preprocess / scan / scatter / sort of the first views — which
    # need neither the texture nor the bucket — run under it.
    ev_state = {"joined": None, "tail_done": None}

    def reset_pipeline():
        """Plain state: the bucket is clear, the next batch forks from the main stream."""
        ev_state["joined"] = ev_state["tail_done"] = None
        if bwd:
            bucket.zero()

    def run_batch(tm=None, nstreams=streams, record=False, final=False, tail_hook=None, **view_kw):
        """One batch of this rank's views + what follows it. ``final``: leave the reduced gradients in the bucket (checksum)."""
        if opt is None:
            invalidate_packed_cache()          # one texture update per step: the packed (6,R,R,4) copy is rebuilt once per batch
        render_views_accumulate(render_fn, g, cams, view_kw.pop("cotangents", cot), views, bg, timer=tm, bucket=bucket, streams=nstreams,
                                backward=bwd, fork_event=ev_state["joined"], render_event=ev_state["tail_done"], **view_kw)
        if not bwd:
            return
        joined = torch.cuda.Event()
        joined.record()
        if record:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            finish_step()
            ev[1].record()
            fin_ev.append(ev)
        else:
            finish_step()
        if tail_hook is not None:
            tail_hook()
        if final:
            ev_state["joined"] = ev_state["tail_done"] = None
            return
        if opt is not None:
            ensure_packed_texture(g.get_texture)       # no-op on one GPU (TextureAdam emits the packed copy itself)
        bucket.zero()                                  # for the next batch
        tail = torch.cuda.Event()
        tail.record()
        ev_state["joined"], ev_state["tail_done"] = (joined, tail) if opt is not None else (None, None)

    step = run_batch
    reset_pipeline()
    for _ in range(args.warmup):
        step()
    sync_all()
    stats = last_stats()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step(record=True, final=(i == args.steps - 1))
    e1.record()
    sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    finish_ms = sum(a.elapsed_time(b) for a, b in fin_ev) / max(1, len(fin_ev))
    if world > 1:
        tms = torch.tensor([ms, wall_ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, wall_ms = float(tms[0].item()), float(tms[1].item())
    clock_rec = clocks.stop() if rank == 0 else None
    value = args.views * args.steps / (ms / 1e3)

    # ---- checksum of the all-reduced gradient bucket: the same number at every N (up to the order of the atomics) ----
    checksum = None
    if bwd:
        rest = [bucket.flat[a:b].double() for a, b in bucket.ranges_without(("texture",))]
        checksum = {"gaussian_grads_l1": float(sum(r.abs().sum() for r in rest).item()),
                    "gaussian_grads_sum": float(sum(r.sum() for r in rest).item()),
                    "what": f"all-reduced per-Gaussian gradients of the last timed step ({args.views} views)"}
        if opt is None:
            tg = bucket.grads()["texture"].double()
            checksum.update(texture_grad_l1=float(tg.abs().sum().item()), texture_grad_sum=float(tg.sum().item()))
        else:
            tx = g.get_texture.detach().double()
            checksum.update(texture_l1=float(tx.abs().sum().item()), texture_sum=float(tx.sum().item()), texture_l1_initial=tex0_l1,
                            optimizer_steps=args.warmup + args.steps,
                            texture_what="the texture after warmup + steps Adam updates: the same on every rank and for every N "
                                         "(up to the order of the gradient atomics)")

    # ---- per-kernel durations: ONE stream, so that no other view's kernels share the SMs with the kernel being timed ----
    stage_ms, timing_note = {}, None
    if not args.no_stage_pass:
        nsteps = max(1, min(args.steps, 4))
        timer = StageTimer(capacity=max(1, len(views)) * nsteps, device=dev)
        reset_pipeline()
        for _ in range(nsteps):
            step(timer, 1)
        sync_all()
        stage_ms = timer.summary()
        timing_note = (f"CUDA events recorded by the library around every kernel, mean over {nsteps} single-stream steps run right after the "
                       f"timed region ({streams} streams there: concurrent views would share the SMs with the kernel being timed)")

    # ---- e2e: the training-shaped step, driven from host buffers -------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            reset_pipeline()
            e2e = run_e2e(args, wl, g, cams, bg, bucket, views, dev, world, sync_all, render_fn, streams, run_batch)
        except Exception as e:              # keep the device-resident measurement; the line then says why e2e is missing
            if world > 1:
                raise                       # a rank that drops out of the collectives would hang the others
            e2e = {"value": None, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(e)[:300]}

    # ---- unique texels touched by one view (for the algorithmic byte count) ---------------------
    if bwd:
        sync_all()
        bucket.zero()
        render_views_accumulate(uv_tex_render, g, cams, [torch.ones_like(c) for c in cot], views[:1] or [0], bg, bucket=bucket)
        U = int((bucket.grads()["texture"].abs().sum(dim=-1) > 0).sum().item())
        bucket.zero()
    else:
        g2 = g.to(requires_grad=True)
        pkg = uv_tex_render(cams[(views[:1] or [0])[0] % len(cams)], g2, None, bg)
        torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], [torch.ones_like(c) for c in cot])
        U = int((g2.get_texture.grad.abs().sum(dim=-1) > 0).sum().item())
        del g2, pkg

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    roofline = roofline_report(wl, args.views, world, value, stage_ms, stats, U, clock_rec, sms, timing_note)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        if full_affinity is not None:
            os.sched_setaffinity(0, full_affinity)      # the CPU leg gets every core again, like the --impl reference run
        try:
            vs = [cpu_oracle_views_per_s(wl, args.cpu_threads) for _ in range(3)]       # a few seconds each; first call builds the scene
            v, desc, threads, _ = max(vs, key=lambda r: r[0])
            cpu = {"value": v, "unit": "views/s", "cores": threads, "kind": "port", "sample": desc + "; best of 3 views"}
        except Exception as e:                                        # never lose the GPU measurement to the CPU leg
            cpu = {"value": None, "unit": "views/s", "cores": 0, "kind": "port", "sample": ("unavailable: " + repr(e))[:300]}

    kernels_per_view = (KERNELS_PER_VIEW if bwd else 6)
    line = {"metric": metric_name(wl), "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, args.views),
            "impl_notes": {"streams_per_rank": streams,
                           "parallelism": ((f"dp{world} (views sharded; per step: fused texture-gradient reduce + Adam + broadcast kernel over NVLink "
                                            f"[{'multimem / NVLS' if opt.multicast else 'peer loads / stores'}], NCCL all-reduce of the other "
                                            f"{sum(b - a for a, b in bucket.ranges_without(('texture',))) * 4 / 1e6:.0f} MB)") if (world > 1 and fused_dp)
                                           else (f"dp{world} (views sharded, 1 all-reduce of {bucket.nbytes / 1e6:.0f} MB/step)" if (world > 1 and bwd)
                                                 else (f"dp{world} (views sharded)" if world > 1 else "single GPU"))),
                           "fallback": fused_note,
                           "optimizer_in_step": (None if not bwd else ("none (--no-optimizer)" if opt is None else "texture Adam (lr 0.0025, eps 1e-15)")),
                           "l2_policy": "inputs larger than L2 (texture %d MB + records %d MB, a different camera every view)"
                                        % (6 * wl.tex_res ** 2 * 12 // 1000000, wl.n_gaussians * 128 // 1000000),
                           "render_fn": render_fn.__name__, "host_wall_ms_per_step": wall_ms / args.steps, "numa_node": numa_node,
                           "reduce_and_optimizer_ms_per_step": finish_ms},
            # our kernels in the timed region, per step: 8 (6 forward-only) per view + the texture step (one GPU: the fused Adam
            # kernel, which also emits the packed copy; N GPUs: the fused reduce + Adam + broadcast kernel and the repack;
            # --no-optimizer: the repack alone)
            "clocks": clock_rec, "e2e": e2e,
            "gpu_launches": (kernels_per_view * len(views) + (0 if not bwd else (2 if fused_dp else 1))) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "grad_checksum": checksum}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def make_supervision(wl, n_views: int, seed: int = 5):
    """Per-view supervision of the training-shaped step in the formats a loader would hand over: uint8 RGB image,
    uint8 alpha mask, int8 normal map (x127) — 7 bytes per pixel and view, pinned. Synthetic (band-limited noise)."""
    gen = torch.Generator().manual_seed(seed)
    H, W = wl.height, wl.width
    out = []
    for _ in range(n_views):
        img = (torch.rand(3, H // 8 + 1, W // 8 + 1, generator=gen)[None])
        img = torch.nn.functional.interpolate(img, size=(H, W), mode="bilinear", align_corners=False)[0]
        nrm = torch.nn.functional.normalize(torch.randn(3, H // 8 + 1, W // 8 + 1, generator=gen), dim=0)[None]
        nrm = torch.nn.functional.normalize(torch.nn.functional.interpolate(nrm, size=(H, W), mode="bilinear", align_corners=False)[0], dim=0)
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
        mask = ((xx * W / H) ** 2 + yy ** 2 < 0.95)[None]
        out.append(((img * 255).round().to(torch.uint8).contiguous().pin_memory(),
                    (mask.to(torch.uint8) * 255).contiguous().pin_memory(),
                    (nrm * 127).round().to(torch.int8).contiguous().pin_memory()))
    return out


def run_e2e(args, wl, g, cams, bg, bucket, views, dev, world, sync_all, render_fn, streams, run_batch):
    """The step a user of the reference runs, through the public operators, with HOST inputs (train.py:147-149,
    models/texture_gaussian3d.py:315-368 with the losses configs/texture_gaussian3d.yaml:77-88 enables): per view the
    ground-truth image (uint8), the alpha mask (uint8) and the normal prior (int8) are copied from pinned host memory on
    a copy stream (one slot per render stream + 1, so the copy of view i+1 runs under the render of view i), decoded on
    the device, the render's losses (1-l)*L1 + l*(1-SSIM), L1(alpha), masked normal loss and bilateral normal smoothness
    come from the fused loss kernels (texture_gs_b200.losses), backward, gradients into the bucket, all-reduce, and
    the per-step loss is read back to the host. Forward-only workloads (retexture.py) read the rendered image back
    instead. All copies are inside the timed region; the time is max(CUDA events, host wall clock)."""
    import torch.distributed as dist
    from texture_gs_b200 import invalidate_packed_cache
    from texture_gs_b200.dist import render_views_accumulate
    from texture_gs_b200.losses import training_loss
    bwd = wl.backward
    H, W = wl.height, wl.width
    lam, lam_alpha, lam_norm, lam_nsm = 0.2, 1.0, 0.1, 0.5          # configs/texture_gaussian3d.yaml:77-88
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    nv = len(views)
    result_host = torch.zeros(1).pin_memory()

    if not bwd:
        # retexture.py-shaped: render, hand the 8-bit frame(s) to the host (it writes PNGs / feeds the viewer)
        n_img = wl.renders_per_view
        frames = [torch.empty(n_img, 3, H, W, dtype=torch.uint8).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        h2d, d2h = (16 + 16 + 3) * 4 * nv, n_img * 3 * H * W * nv

        def step():
            invalidate_packed_cache()
            for i, v in enumerate(views):
                with torch.no_grad():
                    pkg = render_fn(cams[v % len(cams)], g, None, bg)
                    imgs = [pkg["render"]] + ([pkg["render_no_sh"]] if n_img == 2 else [])
                    q = torch.stack([(im.clamp(0, 1) * 255).to(torch.uint8) for im in imgs])
                done[i & 1].synchronize()                      # the pinned frame of two views ago has left the device
                frames[i & 1].copy_(q, non_blocking=True)
                done[i & 1].record(main)
    else:
        sup = make_supervision(wl, min(nv, 4) or 1)
        nslots = streams + 1
        slots = [[torch.empty_like(t, device=dev) for t in sup[0]] for _ in range(nslots)]
        ready = [torch.cuda.Event() for _ in range(nslots)]
        free = [torch.cuda.Event() for _ in range(nslots)]
        h2d, d2h = (sum(t.numel() * t.element_size() for t in sup[0]) + (16 + 16 + 3) * 4) * nv, 4
        total = torch.zeros(streams, device=dev)                       # one partial loss per render stream

        def upload(i):
            s = i % nslots
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[s])
                for d, h in zip(slots[s], sup[i % len(sup)]):
                    d.copy_(h, non_blocking=True)
                ready[s].record(copy_stream)

        cur = [0]                                                      # slot of the view being rendered (host order)

        def before_view(v, i):
            cur[0] = i
            if i + 1 < nv:
                upload(i + 1)                                          # next view's supervision under this view's render
            torch.cuda.current_stream(dev).wait_event(ready[i % nslots])

        def loss_fn(pkg, v):
            k = cur[0]
            u8, m8, n8 = slots[k % nslots]
            gt, gt_alpha, gt_norm = u8 * (1.0 / 255.0), m8 * (1.0 / 255.0), n8 * (1.0 / 127.0)      # one decode kernel each
            loss, _parts = training_loss(pkg["render"], pkg["alpha"], pkg["norm"], gt, gt_alpha, gt_norm,
                                         lam, lam_alpha, lam_norm, lam_nsm)
            total[k % streams] += loss.detach()
            return loss

        def after_view(v, i):
            free[i % nslots].record(torch.cuda.current_stream(dev))

        def read_back():                       # tail of the batch, before the next batch's renders may add to ``total`` again
            result_host.copy_(total.sum().reshape(1), non_blocking=True)
            total.zero_()

        def step():
            upload(0)          # slot 0 was released by the last view that used it (after_view); nothing to wait for on main
            run_batch(cotangents=None, loss_fn=loss_fn, before_view=before_view, after_view=after_view, tail_hook=read_back)

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
    wall_ms = (time.perf_counter() - t0) * 1e3
    ev_ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ev_ms, wall_ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ev_ms, wall_ms = float(tms[0].item()), float(tms[1].item())
    ms = max(ev_ms, wall_ms)
    return {"value": args.views * args.steps / (ms / 1e3), "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": ms / args.steps, "event_ms_per_step": ev_ms / args.steps, "host_wall_ms_per_step": wall_ms / args.steps,
            "result": float(result_host.item()),
            "what": ("render + uint8 frame(s) read back per view (retexture.py)" if not bwd else
                     "uint8 image + uint8 mask + int8 normals H2D per view -> render -> fused losses (photometric, alpha, normal, "
                     "normal smoothness) -> backward -> gradient reduction (+ texture Adam) -> loss D2H per step")}


if __name__ == "__main__":
    main()
