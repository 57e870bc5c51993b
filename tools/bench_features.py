"""Developer measurement (GPU box): achieved HBM bandwidth of the feature / min-max / window kernels (next row 2)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optistate_b200 import kf_batch  # noqa: E402
from optistate_b200.features import assemble_features, min_max, normalized_windows  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best, out


def main():
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    S, T, N = 1024, 1000, 8192
    st = make_streams(range(S), T)
    out = {"hbm_peak_gbs": peak}
    for name, dt, esz in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
        dev = {k: torch.from_numpy(v).to("cuda", dt) for k, v in st.items()}
        res = kf_batch(dev["imu"], dev["p"], dev["dp"], dev["contact"], dev["f"], n_traj=N, dtype=dt, outputs=("x_steps", "p_world_steps"))
        s, rows = timed(lambda: assemble_features(res.x_steps, res.p_world_steps, dev["imu"], dev["f"], dev["dp"], dev["imu_acc"]))
        b = N * T * (24 * esz + 60 * esz)  # unique reads of the per-trajectory arrays + row writes (shared streams stay in L2)
        out[f"assemble_{name}_gbs"] = b / s / 1e9
        flat = rows.reshape(-1, 60)
        s, (lo, hi) = timed(lambda: min_max(flat))
        out[f"minmax_{name}_gbs"] = flat.numel() * esz / s / 1e9
        latent = torch.rand((N * T, 128), dtype=torch.float32, device="cuda")
        s, win = timed(lambda: normalized_windows(flat, lo, hi, latent, n_groups=N, seq_len=10), reps=3)
        b = win.numel() * 4 + flat.numel() * esz + latent.numel() * 4
        out[f"windows_{name}_gbs"] = b / s / 1e9
        out[f"windows_{name}_ms"] = s * 1e3
        del win, latent, rows, flat, res
        torch.cuda.empty_cache()
    for k, v in out.items():
        print(f"{k}: {v:.1f}" + (f"  ({100*v/peak:.0f}% of measured HBM copy peak)" if k.endswith("gbs") and k != "hbm_peak_gbs" else ""))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/features_bw.json", "w"), indent=1)


if __name__ == "__main__":
    main()
