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
    N, M, V, K, U, P, R = 500_000, 15, 485_390, 1_015_518, 22_962_195, 1920 * 1080, 2048
    b = bench.algorithmic_bytes(N, M, V, K, U, P, R, views_per_step=32)
    assert b["render_fwd"] == K * 132 + U * 12 + P * 40                      # list entry + 128-B record, texels, outputs
    assert b["render_bwd"] == P * 40 + K * 132 + 3 * U * 12 + 2 * V * 80
    assert b["preprocess_fwd"] == N * (92 + 12 * M) + V * 140 + 4 * K
    assert b["bwd_clear"] == N * 96 + 6 * R * R * 16 // 32                   # texture-gradient fill charged once per step
    assert abs(sum(b.values()) / 1e6 - 2215.5) < 0.1                          # the per-view figure DESIGN.md §5 quotes


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
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "cfg0_10k_256", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "views/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "pairs" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg0_10k_256" and d["gpu_launches"] == 0


def test_product_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_roofline_report_from_round1_measurements():
    """The reporting arithmetic fed with the stage times and counters of the committed round-1 run reproduces the
    per-kernel figures of profiles/r1_bench_1gpu.json; an empty stage table (no profiling slots) still yields a report."""
    from collections import namedtuple
    from texture_gs_b200.scene import WORKLOADS
    ref = json.loads((ROOT / "profiles" / "r1_bench_1gpu.json").read_text())
    rl = ref["roofline"]
    Stats = namedtuple("Stats", "num_visible num_pairs max_tile_len")
    stats = Stats(rl["counts"]["V"], rl["counts"]["K"], rl["counts"]["max_tile_len"])
    stage_ms = {k: v["ms"] for k, v in rl["per_kernel"].items()}
    stage_ms["forward_total"] = 1.0                      # keys without a byte formula are ignored
    r = bench.roofline_report(WORKLOADS["cfg2_500k_1080p"], 32, 1, ref["value"], stage_ms, stats, rl["counts"]["U"], ref["clocks"], 148)
    assert r["kernel"] == "render_bwd" and r["bound"] == "hbm" and r["unit"] == "GB/s"
    peak, _ = bench.measured_peaks()
    for k, v in rl["per_kernel"].items():
        assert abs(r["per_kernel"][k]["alg_mb"] - v["alg_mb"]) < 0.02, k
        assert abs(r["per_kernel"][k]["gbs"] - v["gbs"]) <= 0.002 * v["gbs"] + 0.2, k
    assert abs(r["achieved"] - 755.0) < 1.0 and abs(r["frac"] - 755.0 / peak) < 1e-3
    assert r["traffic"] == json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())["render_bwd"]
    assert r["algorithmic_bytes_per_launch"] == 1121293796
    assert abs(r["whole_path"]["alg_mb_per_view"] - 2215.5) < 0.1
    iss = r["issue"]
    assert iss["warp_instructions_per_launch"] == 989917246 and 0.5 < iss["frac"] < 0.65      # ncu measured 61 % issue-slot use
    assert abs(iss["peak_ginst_s"] - 148 * 4 * 1.965) < 0.1
    json.dumps(r)
    r0 = bench.roofline_report(WORKLOADS["cfg2_500k_1080p"], 32, 1, ref["value"], {}, stats, rl["counts"]["U"], None, 148)
    assert r0["kernel"] is None and r0["per_kernel"] == {} and r0["achieved"] > 0
    json.dumps(r0)
