"""Developer helper (GPU box): trajectory-steps/s of single bench-shaped launches (1 M trajectories, per-member Q/R).
usage: python tools/rate.py f64:summary f64:x_final f32:summary f64:x_steps ...   (env CASE_N, CASE_T, CASE_STRUCTURE=auto|full)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optistate_b200 import kf_batch  # noqa: E402
from optistate_b200 import _native as nv  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402

N, T, S = int(os.environ.get("CASE_N", 1 << 20)), int(os.environ.get("CASE_T", 300)), 1024
st = make_streams(range(S), T)
rng = np.random.default_rng(0)
q64, r64 = 0.01 * 10 ** rng.uniform(-0.5, 0.5, (12, N)), 0.01 * 10 ** rng.uniform(-0.5, 0.5, (10, N))
for spec in sys.argv[1:]:
    name, out = spec.split(":")
    dt = torch.float64 if name == "f64" else torch.float32
    dev = {k: torch.from_numpy(v).to("cuda", dt) for k, v in st.items()}
    n = N if out != "x_steps" else min(N, 1 << 18)
    kw = dict(Q=torch.from_numpy(q64[:, :n].copy()).to("cuda", dt), R=torch.from_numpy(r64[:, :n].copy()).to("cuda", dt),
              q_kind=nv.MAT_DIAG_PER, r_kind=nv.MAT_DIAG_PER, n_traj=n, dtype=dt, outputs=(out,), structure=os.environ.get("CASE_STRUCTURE", "auto"))
    if out == "mpc":  # predict_mpc covariance model, summary output
        out = "summary"
        kw.update(outputs=("summary",), cov_model="mpc", body_ref=dev["truth"])
    if out == "summary":
        kw.update(truth=dev["truth"], nominal=dev["truth"] * 0.5)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = kf_batch(dev["imu"], dev["p"], dev["dp"], dev["contact"], dev["f"], **kw)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{spec}: {n * T / best / 1e-3:.4e} steps/s  ({best:.2f} ms, {n} x {T})", flush=True)
    del dev, kw, res
