"""TEST INFRASTRUCTURE ONLY - loader for the UNMODIFIED reference Kalman filter.

Imports /root/reference/kalman_filter/kalman_filter.py::Kalman_Filter exactly as shipped, by
(1) planting a stub `casadi` module (the reference's misc/force_controller.py:12 does
`from casadi import *`; casadi is a third-party wheel that is not installed offline) and
(2) replacing `StanceController` (the CasADi/qpOASES MPC, force_controller.py:15-225, out of
scope) with a no-op class *before* kalman_filter.py binds the name (kalman_filter.py:4).
After that `get_odom`, `set_measurements`, `predict`, `update` and `next_state` run unmodified
on NumPy.  Nothing from the reference is copied into this repository.

The reference tree exists only in the build container; on the GPU box the same files are found in
oracle/_ref/ (verbatim copies staged by oracle/make_ref.py at build time, git-ignored).  This module
is used by `oracle/gen_golden.py` (mints tests/golden/*.npz), by the CPU-side tests that cross-check
the restatements in this directory against the live reference (they skip when neither tree exists),
and by bench.py's `cpu_baseline` leg, which times the reference class itself on the box's host cores.
Product code must never import it.
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

def _default_root() -> str:
    from . import make_ref

    return make_ref.root() or "/root/reference"


REF_ROOT = os.environ.get("OPTISTATE_REF") or _default_root()
_CASADI_NAMES = ["casadi", "vertcat", "horzcat", "mtimes", "if_else", "cos", "sin", "tan", "transpose", "inv", "skew"]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "kalman_filter", "kalman_filter.py"))


def load():
    """Returns (Kalman_Filter class, force_controller module, settings module) of the reference."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    if "casadi" not in sys.modules:
        cas = types.ModuleType("casadi")
        for n in _CASADI_NAMES:
            setattr(cas, n, MagicMock(name=n))
        cas.__all__ = list(_CASADI_NAMES)
        sys.modules["casadi"] = cas
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import misc.force_controller as fc  # noqa: E402  (reference module)

    fc.StanceController = type("NoMPC", (), {"__init__": lambda self, *a, **k: None})
    import settings as ref_settings  # noqa: E402
    from kalman_filter.kalman_filter import Kalman_Filter  # noqa: E402

    return Kalman_Filter, fc, ref_settings


def run_reference(stream, x0=None, P0=None, Q=None, R=None, p_checkpoint_every=0, mode="predict"):
    """Runs the reference class over one base stream (dict of [T, C] arrays).

    Per step: get_odom -> set_measurements -> predict(p, f) -> update  (SURVEY 3.2), or with
    mode="mpc_cov" the non-QP lines of predict_mpc (kalman_filter.py:153-161) with the supplied f
    and body_ref = stream["body_ref"][t].
    Returns dict(x [T,12], x_model [T,12], p_trace [T], k_gain [T], z [T,10], p_world [T,12],
                 P_final [12,12], P_ckpt {step: [12,12]}, K_last [12,10]).
    """
    KF_cls, fc, ref_settings = load()
    kf = KF_cls()
    # never `kf.x[:] = ...`: x aliases INITIAL_PARAMS.STARTING_STATE (kalman_filter.py:10)
    kf.x = (ref_settings.INITIAL_PARAMS.STARTING_STATE if x0 is None else np.asarray(x0, float)).reshape(12, 1).copy()
    if Q is not None:
        kf.Q = np.array(Q, dtype=float)
    if R is not None:
        kf.R = np.array(R, dtype=float)
    kf.P = np.array(kf.Q if P0 is None else P0, dtype=float).copy()
    kf.F = kf.F.astype(float).copy()

    T = stream["imu"].shape[0]
    out = {
        "x": np.empty((T, 12)), "x_model": np.empty((T, 12)), "p_trace": np.empty(T), "k_gain": np.empty(T),
        "z": np.empty((T, 10)), "p_world": np.empty((T, 12)), "P_ckpt": {},
    }
    for t in range(T):
        imu = stream["imu"][t].reshape(6, 1).copy()
        p = stream["p"][t].reshape(12, 1).copy()
        dp = stream["dp"][t].reshape(12, 1).copy()
        contact = stream["contact"][t].reshape(4, 1).copy()
        f = stream["f"][t].reshape(12, 1).copy()
        kf.set_measurements(imu, kf.get_odom(p, dp, contact, imu))
        if mode == "predict":
            kf.predict(p, f)
        elif mode == "mpc_cov":
            body_ref = stream["body_ref"][t].reshape(12, 1)
            Rm = kf.rotation_matrix_body_world(body_ref[0], body_ref[1], body_ref[2])
            kf.F[0:3, 6:9] = np.transpose(Rm)
            kf.F_d = np.exp(kf.dt * kf.F)
            kf.P = np.matmul(np.matmul(kf.F_d, kf.P), np.transpose(kf.F_d)) + kf.Q
            kf.x = fc.next_state(kf.x, p, f, kf.dt)
            kf.x_model = kf.x.copy()
        else:
            raise ValueError(mode)
        out["x_model"][t] = kf.x.reshape(12)
        kf.update()
        out["x"][t] = kf.x.reshape(12)
        out["p_trace"][t] = kf.P_trace
        out["k_gain"][t] = kf.K_gain
        out["z"][t] = kf.z.reshape(10)
        out["p_world"][t] = p.reshape(12)
        if p_checkpoint_every and (t + 1) % p_checkpoint_every == 0:
            out["P_ckpt"][t + 1] = kf.P.copy()
    out["P_final"] = kf.P.copy()
    out["K_last"] = kf.K.copy()
    return out
