"""Developer helper (GPU box): run a few single kf_batch launches so that ncu can capture them.
usage: python tools/ncu_case.py f64:tma:x_final f32:direct:summary f64:tma-mpc:x_steps ...   (env CASE_N, CASE_T, CASE_S)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optistate_b200 import kf_batch  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402

N, T, S = int(os.environ.get("CASE_N", 606208)), int(os.environ.get("CASE_T", 100)), int(os.environ.get("CASE_S", 1024))
st = make_streams(range(S), T)
for spec in sys.argv[1:]:
    dt_name, mode, out = spec.split(":")
    dt = torch.float64 if dt_name == "f64" else torch.float32
    dev = {k: torch.from_numpy(v).to("cuda", dt) for k, v in st.items()}
    kw = {}
    if mode.endswith("-mpc"):  # predict_mpc covariance model
        mode = mode[:-4]
        kw.update(cov_model="mpc", body_ref=dev["truth"])
    if mode == "joint":
        kw["algo"] = "joint"
    if mode == "direct":
        kw["stream_index"] = torch.arange(N, dtype=torch.int32, device="cuda") % S
    if out.startswith("summary"):
        kw["truth"] = dev["truth"]
        if out == "summary2":
            kw["nominal"] = dev["truth"] * 0.5
        out = "summary"
    for _ in range(2):
        res = kf_batch(dev["imu"], dev["p"], dev["dp"], dev["contact"], dev["f"], n_traj=N, dtype=dt, outputs=(out,), **kw)
    torch.cuda.synchronize()
    print(spec, res.algo, "ok", flush=True)
