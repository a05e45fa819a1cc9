"""Host-side pieces of bench.py that do not need a GPU: the byte accounting of the roofline, the clock-sample parser,
the CPU arm (``--impl reference``: the oracle on host cores) on the small BASELINE config, and the refusal of the
product arm to run without CUDA."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_algorithmic_bytes_follow_the_survey_formula():
    """SURVEY.md §8(d), with the survey's per-unit sizes (not this implementation's 128-byte record / 16-byte texel)."""
    N, M, V, K, U, P, R = 500_000, 15, 485_390, 1_015_518, 22_962_195, 1920 * 1080, 2048
    b = bench.algorithmic_bytes(N, M, V, K, U, P, R, views_per_step=32)
    fwd = N * (92 + 12 * M) + V * 64 + K * 24 + K * 116 + U * 12 + P * 40
    bwd = P * 40 + K * 116 + U * 12 + 6 * R * R * 12 // 32 + 2 * U * 12 + 2 * V * 68 + N * (92 + 12 * M) + N * (68 + 12 * M)
    assert b["render_fwd"] == K * 116 + U * 12 + P * 40
    assert b["render_bwd"] == P * 40 + K * 116 + 3 * U * 12 + 2 * V * 68
    assert sum(b[k] for k in ("preprocess_fwd", "scan_tiles", "scatter_pairs", "sort_tiles", "render_fwd")) == fwd
    assert sum(b[k] for k in ("render_bwd", "bwd_clear", "preprocess_bwd")) == bwd
    f = bench.algorithmic_bytes(N, M, V, K, U, P, R, views_per_step=32, backward=False)
    assert sum(f.values()) == fwd and "render_bwd" not in f


def test_clock_sampler_parses_nvidia_smi_lines_and_reports_throttle_reasons():
    c = bench.ClockSampler(0)
    c.proc = type("P", (), {"terminate": lambda s: None, "wait": lambda s, timeout=None: 0, "kill": lambda s: None})()
    c.lines = ["1965, 1965, 700.1, Not Active, Not Active, Not Active, Not Active",
               "1950, 1965, 740.0, Not Active, Not Active, Not Active, Active",
               "garbage", "1800, 1965, [N/A], Not Active, Active, Not Active, Not Active"]
    r = c.stop()
    assert r["sm_mhz"] == 1950.0 and r["sm_max_mhz"] == 1965.0 and r["samples"] == 3 and r["power_w_max"] == 740.0
    assert r["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]     # a sample without a power reading still counts


def test_measured_peaks_file_or_documented_fallback():
    peak, src = bench.measured_peaks()
    assert peak > 1000 and (src.startswith("measured") or src.startswith("fallback"))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """Under torchrun every rank gets OMP_NUM_THREADS=1: the CPU arm must not inherit that (round 1 did, and its
    multi-GPU baselines were single-threaded); it runs every requested step and reports a measured ms_per_step."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    env.pop("TEXGS_CPU_THREADS", None)
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "cfg0_10k_256", "--steps", "3",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "views/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] == 1
    assert abs(d["ms_per_step"] - 1000.0 / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["cpu_baseline"]["kind"] == "port" and "pairs" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))             # not the injected OMP_NUM_THREADS
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    # the config object is the SAME in both arms: built by one function from the workload alone
    from texture_gs_b200.scene import WORKLOADS
    assert d["config"] == bench.workload_config(WORKLOADS["cfg0_10k_256"], 32)
    assert set(d["config"]) == {"workload", "gaussians", "width", "height", "tex_res", "views_per_step", "sh_degree", "pass", "renders_per_view"}
    r2 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "cfg0_10k_256", "--steps", "1",
                         "--warmup", "0", "--cpu-threads", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert json.loads(r2.stdout.strip().splitlines()[-1])["cpu_baseline"]["cores"] == 2


def test_product_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_roofline_report_from_round1_measurements():
    """The reporting arithmetic fed with the stage times and counters of the committed round-1 run: achieved = SURVEY
    bytes / measured duration for every kernel; an empty stage table (no profiling pass) still yields a report."""
    from collections import namedtuple
    from texture_gs_b200.scene import WORKLOADS
    ref = json.loads((ROOT / "profiles" / "r1_bench_1gpu.json").read_text())
    rl = ref["roofline"]
    Stats = namedtuple("Stats", "num_visible num_pairs max_tile_len")
    c = rl["counts"]
    stats = Stats(c["V"], c["K"], c["max_tile_len"])
    stage_ms = {k: v["ms"] for k, v in rl["per_kernel"].items()}
    stage_ms["forward_total"] = 1.0                      # keys without a byte formula are ignored
    wl = WORKLOADS["cfg2_500k_1080p"]
    r = bench.roofline_report(wl, 32, 1, ref["value"], stage_ms, stats, c["U"], ref["clocks"], 148, "note")
    assert r["kernel"] == "render_bwd" and r["bound"] == "hbm" and r["unit"] == "GB/s" and r["timing"] == "note"
    peak, _ = bench.measured_peaks()
    ab = bench.algorithmic_bytes(wl.n_gaussians, 15, c["V"], c["K"], c["U"], wl.width * wl.height, wl.tex_res, 32)
    for k, v in rl["per_kernel"].items():
        if ab[k] == 0:
            continue
        assert abs(r["per_kernel"][k]["gbs"] - ab[k] / (v["ms"] * 1e-3) / 1e9) < 0.1, k
    # the judge's recompute of round 1: 40P + 116K + 36U + 136V = 1093 MB / 1.4838 ms -> 0.112 of the measured peak
    assert abs(r["algorithmic_bytes_per_launch"] / 1e6 - 1093.4) < 0.5
    assert abs(r["frac"] - r["algorithmic_bytes_per_launch"] / (stage_ms["render_bwd"] * 1e-3) / 1e9 / peak) < 1e-3
    assert r["traffic"] == json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())["render_bwd"]
    iss = r["issue"]
    want = json.loads((ROOT / "profiles" / "ncu_issue.json").read_text())["render_bwd"]["warp_instructions"]
    assert iss["warp_instructions_per_launch"] == want and 0.5 < iss["frac"] < 0.65      # ncu measured 59-61 % issue-slot use
    assert abs(iss["peak_ginst_s"] - 148 * 4 * 1.965) < 0.1
    json.dumps(r)
    r0 = bench.roofline_report(wl, 32, 1, ref["value"], {}, stats, c["U"], None, 148)
    assert r0["kernel"] is None and r0["per_kernel"] == {} and r0["achieved"] > 0
    json.dumps(r0)
