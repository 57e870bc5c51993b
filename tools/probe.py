"""Developer probe (GPU box): kernel-only throughput of the filter kernels at a few sizes + the FMA peaks."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optistate_b200 import fma_peak, kf_batch  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    out = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        fl, secs = fma_peak(dt, 1 << 18)
        out[f"fma_peak_{name}_tflops"] = fl / 1e12
        print(f"fma peak {name}: {fl/1e12:.2f} TFLOP/s ({secs*1e3:.2f} ms)", flush=True)
    S, T = 1024, int(os.environ.get("PROBE_T", 200))
    st = make_streams(range(S), T)
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        dev = {k: torch.from_numpy(v).to("cuda", dt) for k, v in st.items()}
        for N in [int(x) for x in os.environ.get("PROBE_N", "1024,151552,606208,1212416").split(",")]:
            for algo in ("sequential", "joint"):
                if algo == "joint" and N > 160000:
                    continue
                for outs in (("x_final",), ("summary",), ("x_steps",)) if algo == "sequential" else (("x_final",),):
                    if outs == ("x_steps",) and N * T * 12 * (8 if name == "f64" else 4) > 40e9:
                        continue
                    kw = dict(truth=dev["truth"]) if outs == ("summary",) else {}
                    fn = lambda: kf_batch(dev["imu"], dev["p"], dev["dp"], dev["contact"], dev["f"], n_traj=N, dtype=dt, algo=algo, outputs=outs, **kw)  # noqa: E731
                    s = timed(fn)
                    rate = N * T / s
                    out[f"{algo}_{name}_N{N}_{outs[0]}"] = rate
                    print(f"{algo:10s} {name} N={N:8d} T={T} out={outs[0]:8s}: {s*1e3:9.2f} ms  {rate:.3e} steps/s  ({rate*6800/1e12:.2f} TF/s @6800)", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
