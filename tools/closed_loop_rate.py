"""Developer helper (GPU box): steps/s of the batched closed loop estimate_state_mpc (MPC from the current estimate + one filter
step, per step).  usage: python tools/closed_loop_rate.py [n_trajectories] [n_steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optistate_b200.mpc import estimate_state_mpc_batch  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
st = {k: torch.from_numpy(v).cuda() for k, v in make_streams(range(min(N, 1024)), T).items()}
rep = max(1, N // min(N, 1024))
st = {k: v.repeat(1, 1, rep) for k, v in st.items()}
ref = torch.zeros((T, 5, 12, st["imu"].shape[2]), dtype=torch.float64, device="cuda")
ref[:, :, 5] = 0.28
from optistate_b200.mpc import WarmStart  # noqa: E402

rounds = int(os.environ.get("WARM_ROUNDS", "0"))
best = float("inf")
for _ in range(5):  # best of five: a process that has just started finds the GPU at idle clocks
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kw = {"warm": False} if rounds < 0 else {}  # default: the dual active-set kernel with its warm start; WARM_ROUNDS=-1: cold
    xs, fs, mst, fst = estimate_state_mpc_batch(st["imu"], st["p"], st["dp"], st["contact"], ref, **kw)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
dt = best
n = xs.shape[2]
it = (mst >> 8).double()
print("per step: warm accepted", [round(float((mst[t] & 8).ne(0).double().mean()), 2) for t in range(min(T, 12))], "ipm iterations", [round(float(it[t].mean()), 1) for t in range(min(T, 12))])
print(f"{n} trajectories x {T} steps: {n * T / dt:.3e} trajectory-steps/s ({dt / T * 1e3:.2f} ms per step); unpolished {int((mst & 2).ne(0).sum())}, "
      f"filter status {int(fst.max())}")
