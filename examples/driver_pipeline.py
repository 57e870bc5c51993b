"""The reference's conversion driver, end to end on the device, on synthetic recordings.

What /root/reference/data_collection/data_conversion_Kalman_to_Training.py does per recording, with the batched calls of
this package in place of the per-step Python loop (the Drive recordings are not available offline, so the streams come
from optistate_b200.synth, shaped like them):

    1. Q / R identification from model-vs-mocap residuals        (:31-109)   -> identify_noise
    2. the filter loop over every step                            (:193-201)  -> kf_batch with the recorded forces (all recordings in one
       launch), or - `--closed-loop`, what the shipped driver calls: KF.estimate_state_mpc, the force MPC solved from the current
       estimate at every step (:199) - estimate_state_mpc_batch
    3. the 60-wide feature rows [x, imu_acc, f, p_world, dp, imu] (:245-254)  -> assemble_features
    4. min-max normalisation and sliding windows for the GRU      (gru_train.py:56-63,108-111,180-192) -> min_max, normalized_windows

    python examples/driver_pipeline.py [n_recordings] [n_steps] [--closed-loop]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optistate_b200 import kf_batch  # noqa: E402
from optistate_b200.features import assemble_features, min_max, normalized_windows  # noqa: E402
from optistate_b200.identify import identify_noise  # noqa: E402
from optistate_b200.mpc import estimate_state_mpc_batch  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402


def run(n_recordings: int = 64, n_steps: int = 4063, seq_len: int = 10, dtype=torch.float64, verbose: bool = False, closed_loop: bool = False):
    st = make_streams(range(n_recordings), n_steps)
    dev = {k: torch.from_numpy(v).to("cuda", dtype) for k, v in st.items()}
    t0 = time.perf_counter()
    # 1. per-recording noise levels (variances of the one-step model / measurement residuals against the label stream)
    q_diag, r_diag, id_status = identify_noise(dev["truth"], dev["imu"], dev["p"], dev["dp"], dev["contact"], dev["f"], dtype=dtype)
    r_diag[0:3] = 1e-4  # the driver's override of the attitude measurement noise (:141-143)
    if closed_loop:
        # 2'. the shipped driver's call: per step, the force MPC from the current estimate, then one filter step with those forces
        #     (reference trajectory: the label stream held over the horizon, as ref_list[i] is in the driver); FP64
        body_ref = dev["truth"].double()[:, None, :, :].expand(-1, 5, -1, -1).contiguous()
        xs, fs, mst, fst, pws = estimate_state_mpc_batch(dev["imu"].double(), dev["p"].double(), dev["dp"].double(), dev["contact"].double(), body_ref,
                                                         Q=q_diag.double().clamp_min(1e-12), R=r_diag.double().clamp_min(1e-12), return_p_world=True)
        rows = assemble_features(xs.to(dtype), pws.to(dtype), dev["imu"], fs.to(dtype), dev["dp"], dev["imu_acc"])
        res = type("ClosedLoop", (), dict(algo="closed loop (force MPC + sequential)", x_steps=xs, status=fst, mpc_status=mst))()
    else:
        # 2. every recording, every step, one launch
        res = kf_batch(dev["imu"], dev["p"], dev["dp"], dev["contact"], dev["f"], Q=q_diag.clamp_min(1e-12), R=r_diag.clamp_min(1e-12),
                       dtype=dtype, outputs=("x_steps", "p_world_steps", "p_trace", "k_gain"))
        # 3. feature rows [N, T, 60]
        rows = assemble_features(res.x_steps, res.p_world_steps, dev["imu"], dev["f"], dev["dp"], dev["imu_acc"])
    # 4. normalise over all rows, windows inside each recording, float32 for the GRU
    flat = rows.reshape(-1, 60)
    lo, hi = min_max(flat)
    windows = normalized_windows(flat, lo, hi, n_groups=n_recordings, seq_len=seq_len)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if verbose:
        print(f"{n_recordings} recordings x {n_steps} steps: algo={res.algo}, {n_recordings * n_steps / dt:.3e} trajectory-steps/s through the "
              f"whole pipeline ({dt * 1e3:.1f} ms); windows {tuple(windows.shape)} {windows.dtype}; "
              f"status: identify {int(id_status.max())}, filter {int(res.status.max())}")
    return res, rows, windows


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if len(args) > 0 else 64
    T = int(args[1]) if len(args) > 1 else 4063
    cl = "--closed-loop" in sys.argv
    run(n, T, verbose=True, closed_loop=cl)   # first call includes CUDA context + library load
    run(n, T, verbose=True, closed_loop=cl)
