"""Q / R identification on the device: the step right BEFORE the filter.

Replaces the `load_Q_R = False` branch of the reference driver
(/root/reference/data_collection/data_conversion_Kalman_to_Training.py:31-109): per recording, the state is reset to the
ground truth at every step, the model is stepped once, the next measurement is formed, and Q / R are the variances of
the model and measurement residuals.  The forces the reference obtains from its MPC inside `predict_mpc` are an input.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nv
from .settings import INITIAL_PARAMS


def identify_noise(gt, imu, p, dp, contact, f, *, n_traj=None, dtype: torch.dtype = torch.float64, stream_index=None,
                   stream_offset: int = 0, alias_last_measurement: bool = False, dt: float = INITIAL_PARAMS.DT_mpc,
                   mass: float = INITIAL_PARAMS.ROBOT_MASS, inertia=None, gravity: float = INITIAL_PARAMS.GRAVITY, device=None):
    """gt [T,12,S] ground-truth states; imu [T,6,S], p, dp, f [T,12,S], contact [T,4,S].
    Returns (q_diag [12,N], r_diag [10,N], status [N]) on the device.  `alias_last_measurement=True` reproduces the
    reference exactly (its list of measurements aliases one array, so every entry is the last measurement)."""
    nv.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    to_dev = lambda a: torch.as_tensor(a).to(device=device, dtype=dtype).contiguous()  # noqa: E731
    tensors = {k: to_dev(v) for k, v in (("gt", gt), ("imu", imu), ("p", p), ("dp", dp), ("contact", contact), ("f", f))}
    T, _, S = tensors["gt"].shape
    N = int(S if n_traj is None else n_traj)
    tensors["q_diag"] = torch.empty((12, N), dtype=dtype, device=device)
    tensors["r_diag"] = torch.empty((10, N), dtype=dtype, device=device)
    tensors["status"] = torch.zeros(N, dtype=torch.int32, device=device)
    if stream_index is not None:
        from .batch import _checked_stream_index

        tensors["stream_index"] = _checked_stream_index(stream_index, N, S, device)
    inertia = np.diag(INITIAL_PARAMS.INERTIA_ROT) if inertia is None else np.asarray(inertia, float).reshape(3)
    consts = dict(dt=float(dt), mass=float(mass), inertia0=float(inertia[0]), inertia1=float(inertia[1]), inertia2=float(inertia[2]),
                  gravity=float(gravity))
    with torch.cuda.device(device):
        rc = nv.ext().kf_identify_noise(nv.F64 if dtype == torch.float64 else nv.F32, N, T, S, int(stream_offset),
                                        int(alias_last_measurement), consts, tensors)
    nv.check(rc, "optistate_kf_identify_noise")
    return tensors["q_diag"], tensors["r_diag"], tensors["status"]
