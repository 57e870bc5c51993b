"""Harness tests (bench.py's measurement contract).  The file name sorts LAST on purpose: a tooling failure here must never
hide a parity test under `pytest -x` (round 1 lost the seven fused-all-gather tests that way)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*extra, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_bench_native_arm_prints_the_contract_line():
    """bench.py on a small workload: every key of the measurement contract is present and self-consistent."""
    d = _bench("--steps", "2", "--warmup", "3", "--traj-per-gpu", "8192", "--T", "60", "--streams", "128", "--no-cpu-baseline", "--no-secondary")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "parity_sample"):
        assert k in d, k
    assert d["metric"] == "kf_trajectory_steps_per_sec" and d["unit"] == "trajectory-steps/s" and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["scaling"] == "weak" and d["config"]["trajectories_total"] == 8192
    assert abs(d["value"] - 8192 * 60 * 2 / (d["ms_per_step"] * 2e-3)) < 1e-6 * d["value"]
    assert d["gpu_launches"] >= 2 * 2  # measurement pre-pass + filter kernel per step
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 52 * 8192 * 8
    r = d["roofline"]
    assert r["bound"] == "fma" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["peak"] > 10
    assert abs(r["frac_algorithmic"] - r["achieved_algorithmic"] / r["peak"]) < 1e-12
    assert r["measured_fma_peak_tflops"] > 0.9 * r["theoretical_peak_at_sampled_clock_tflops"]  # the probe saturates the FP64 pipe
    assert d["status_nonzero_trajectories"] == 0
    # the short region still has its two synchronous clock readings (round 1: a 200 ms poll saw none)
    assert d["clocks"]["sm_mhz"] is not None and d["clocks"]["samples"] >= 2 and d["clocks"]["sm_max_mhz"] >= d["clocks"]["sm_mhz"]
    ps = d["parity_sample"]
    assert ps["ok"] and ps["n"] >= 8 and ps["max_rel_x"] < 1e-9 and ps["max_rel_p"] < 1e-9


def test_bench_strong_scaling_mode_and_full_covariance_structure():
    """--traj-total shards a fixed job (BASELINE configs[3] is --traj-total 16777216); --structure full times the 78-entry kernel."""
    d = _bench("--steps", "1", "--warmup", "3", "--traj-total", "4096", "--T", "50", "--streams", "64", "--no-cpu-baseline", "--no-secondary",
               "--no-e2e", "--structure", "full")
    assert d["scaling"] == "strong" and d["config"]["trajectories_total"] == 4096 and d["parity_sample"]["ok"]
    assert "78" in d["config"]["covariance_structure"]
