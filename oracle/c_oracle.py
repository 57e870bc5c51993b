"""TEST INFRASTRUCTURE ONLY - ctypes front end of oracle/kf_oracle.c (the checker / CPU baseline)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkf_oracle.so")
_lib = None

_DP = C.POINTER(C.c_double)


class KfOracleArgs(C.Structure):
    _fields_ = [
        ("n_traj", C.c_int64), ("n_steps", C.c_int64), ("n_streams", C.c_int64),
        ("cov_model", C.c_int32), ("q_kind", C.c_int32), ("r_kind", C.c_int32),
        ("x0_per_traj", C.c_int32), ("p0_kind", C.c_int32), ("n_threads", C.c_int32),
        ("dt", C.c_double), ("mass", C.c_double), ("inertia", C.c_double * 3), ("gravity", C.c_double),
        ("imu", _DP), ("p", _DP), ("dp", _DP), ("contact", _DP), ("f", _DP), ("body_ref", _DP),
        ("stream_index", C.POINTER(C.c_int32)),
        ("x0", _DP), ("P0", _DP), ("Q", _DP), ("R", _DP),
        ("ckpt_every", C.c_int64),
        ("x_steps", _DP), ("x_model_steps", _DP), ("p_world_steps", _DP), ("z_steps", _DP),
        ("p_trace_steps", _DP), ("k_gain_steps", _DP), ("nis_steps", _DP), ("P_ckpt", _DP),
        ("x_final", _DP), ("P_final", _DP), ("K_final", _DP),
        ("status", C.POINTER(C.c_uint32)),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.kf_oracle_run.argtypes = [C.POINTER(KfOracleArgs)]
        _lib.kf_oracle_run.restype = C.c_int
        _lib.kf_oracle_max_threads.restype = C.c_int
        _lib.kf_oracle_args_size.restype = C.c_size_t
        assert _lib.kf_oracle_args_size() == C.sizeof(KfOracleArgs)
    return _lib


def max_threads() -> int:
    return int(lib().kf_oracle_max_threads())


def _ptr(a, typ=_DP):
    return a.ctypes.data_as(typ) if a is not None else typ()


def run(streams, n_traj=None, *, Q=None, R=None, x0=None, P0=None, stream_index=None, cov_model=0,
        ckpt_every=0, n_threads=0, noise_per_traj=None, want=("x_steps", "p_trace_steps", "k_gain_steps", "x_final", "P_final"),
        dt=0.01, mass=8.8, inertia=(55303643.08 / 10**9, 60119440.34 / 10**9, 105304340.05 / 10**9), gravity=-9.81):
    """Filters n_traj trajectories over base streams laid out [T, C, S] (float64).

    Q: [12,12] shared dense or [12,N] per-trajectory diagonal;  R: [10,10] or [10,N];
    x0: [12] or [12,N];  P0: None (= Q), [12,12] or [144,N].  Returns dict of the arrays named in `want`.
    noise_per_traj: True / False settles how a [12,12] Q (N = 12) or [10,10] R (N = 10) is read; None = shared dense.
    """
    from .kf_numpy import Q_DEFAULT, R_DEFAULT, START

    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    imu, p, dp, contact, f = (f64(streams[k]) for k in ("imu", "p", "dp", "contact", "f"))
    T, _, S = imu.shape
    N = int(S if n_traj is None else n_traj)
    a = KfOracleArgs()
    a.n_traj, a.n_steps, a.n_streams = N, T, S
    a.cov_model, a.n_threads, a.ckpt_every = cov_model, n_threads, ckpt_every
    a.dt, a.mass, a.gravity = dt, mass, gravity
    a.inertia = (C.c_double * 3)(*inertia)
    Q = f64(np.diag(Q_DEFAULT) if Q is None else Q)
    R = f64(np.diag(R_DEFAULT) if R is None else R)
    a.q_kind = 0 if Q.shape == (12, 12) and not (N == 12 and noise_per_traj) else 1
    a.r_kind = 0 if R.shape == (10, 10) and not (N == 10 and noise_per_traj) else 1
    if a.q_kind == 1:
        assert Q.shape == (12, N), Q.shape
    if a.r_kind == 1:
        assert R.shape == (10, N), R.shape
    x0 = f64(START if x0 is None else x0)
    a.x0_per_traj = 0 if x0.ndim == 1 else 1
    if P0 is None:
        a.p0_kind, P0 = 0, None
    else:
        P0 = f64(P0)
        a.p0_kind = 1 if P0.shape == (12, 12) else 2
    keep = [imu, p, dp, contact, f, Q, R, x0, P0]
    a.imu, a.p, a.dp, a.contact, a.f = (_ptr(v) for v in (imu, p, dp, contact, f))
    if cov_model == 1:
        br = f64(streams["body_ref"])
        keep.append(br)
        a.body_ref = _ptr(br)
    if stream_index is not None:
        si = np.ascontiguousarray(stream_index, dtype=np.int32)
        keep.append(si)
        a.stream_index = _ptr(si, C.POINTER(C.c_int32))
    a.x0, a.P0, a.Q, a.R = _ptr(x0), _ptr(P0), _ptr(Q), _ptr(R)
    shapes = {
        "x_steps": (T, 12, N), "x_model_steps": (T, 12, N), "p_world_steps": (T, 12, N), "z_steps": (T, 10, N),
        "p_trace_steps": (T, N), "k_gain_steps": (T, N), "nis_steps": (T, N),
        "P_ckpt": ((T // ckpt_every) if ckpt_every else 0, 144, N),
        "x_final": (12, N), "P_final": (144, N), "K_final": (120, N),
    }
    out = {}
    for name in want:
        out[name] = np.empty(shapes[name], dtype=np.float64)
        setattr(a, name, _ptr(out[name]))
    status = np.zeros(N, dtype=np.uint32)
    a.status = _ptr(status, C.POINTER(C.c_uint32))
    rc = lib().kf_oracle_run(C.byref(a))
    if rc != 0:
        raise RuntimeError(f"kf_oracle_run failed: {rc}")
    out["status"] = status
    del keep
    return out
