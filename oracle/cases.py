"""TEST INFRASTRUCTURE ONLY - the parity cases shared by gen_golden.py (which runs the unmodified
reference on them) and tests/ (which run the oracle restatements and the CUDA path on them).

A case is (stream dict of [T, C] arrays, kwargs for the filter): x0, P0, Q, R (dense), model.
"""
from __future__ import annotations

import os

import numpy as np

from optistate_b200.synth import make_stream

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

Q_DEFAULT = np.diag([0.01, 0.01, 0.01, 0.01, 0.0001, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.0001])
R_DEFAULT = np.diag([0.01] * 10)
START = np.array([0, 0, 0, 0, 0, 0.28, 0, 0, 0, 0, 0, 0], dtype=np.float64)


def q_r_pkl():
    """(Q, R) of /root/reference/data_collection/trajectories/Q_R.pkl with the driver's override
    R[0..2,0..2] = 1e-4 (data_conversion_Kalman_to_Training.py:139-143); values are stored as a fixture."""
    d = np.load(os.path.join(GOLDEN_DIR, "q_r_pkl.npz"))
    R = np.diag(d["r_diag"]).copy()
    R[0, 0] = R[1, 1] = R[2, 2] = 1e-4
    return np.diag(d["q_diag"]).copy(), R


def _spd(rng, n, scale):
    A = rng.standard_normal((n, n))
    return scale * (A @ A.T / n + 0.5 * np.eye(n))


def _contact_patterns(T):
    """Cycles through 1-, 2-, 3- and 4-leg stance patterns (never all-swing)."""
    pats = np.array([[1, 0, 0, 0], [1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 0], [1, 1, 1, 1], [0, 0, 1, 0], [0, 1, 1, 1]], float)
    return pats[(np.arange(T) // 5) % len(pats)]


def build(name: str):
    """Returns (stream, kwargs, subsample) for a named case; subsample = step stride stored in the fixture."""
    kw = dict(x0=START.copy(), P0=None, Q=Q_DEFAULT.copy(), R=R_DEFAULT.copy(), model="predict")
    if name == "cfg1_default_seed0":
        return make_stream(0, 2000), kw, 1
    if name == "default_seed11_10k":
        return make_stream(11, 10000), kw, 10
    if name == "stress_qrpkl_seed3_10k":
        kw["Q"], kw["R"] = q_r_pkl()
        return make_stream(3, 10000), kw, 10
    if name == "edge_zero_attitude_spin":  # trunc(R^T) = I at step 0 with a non-zero body rate
        kw["x0"] = np.array([0, 0, 0, 0.1, -0.2, 0.3, 0.3, -0.2, 0.1, 0.05, -0.05, 0.02])
        return make_stream(21, 64), kw, 1
    if name == "edge_yaw_quarter_turn":  # sin(pi/2) == 1.0 exactly: off-diagonal trunc entries fire
        kw["x0"] = np.array([0, 0, np.pi / 2, 0, 0, 0.28, 0.2, 0.1, -0.3, 0, 0, 0])
        return make_stream(22, 64), kw, 1
    if name == "edge_dense_noise":
        rng = np.random.default_rng(1234)
        kw["Q"], kw["R"], kw["P0"] = _spd(rng, 12, 0.01), _spd(rng, 10, 0.02), _spd(rng, 12, 0.05)
        kw["x0"] = START + 0.05 * rng.standard_normal(12)
        return make_stream(23, 200), kw, 1
    if name == "edge_nonsymmetric_p0":
        rng = np.random.default_rng(4321)
        kw["P0"] = Q_DEFAULT + 1e-3 * 0.01 * rng.standard_normal((12, 12))
        return make_stream(24, 200), kw, 1
    if name == "edge_contact_patterns":
        s = make_stream(25, 210)
        s["contact"] = _contact_patterns(210)
        return s, kw, 1
    if name == "edge_large_angles":
        kw["x0"] = np.array([0.9, -1.2, 2.8, 1.0, -2.0, 0.4, 0.5, -0.4, 0.3, 0.2, 0.1, -0.1])
        s = make_stream(26, 200)
        s["imu"][:, 0:3] += np.array([0.9, -1.2, 2.8])
        return s, kw, 1
    if name == "edge_diag_p0_seed31":  # a diagonal P0 that is NOT Q (the decoupled-group kernels take any diagonal P0), Q_R.pkl noise
        rng = np.random.default_rng(31)
        kw["Q"], kw["R"] = q_r_pkl()
        kw["P0"] = np.diag(rng.uniform(1e-3, 0.5, 12))
        kw["x0"] = START + 0.02 * rng.standard_normal(12)
        return make_stream(31, 300), kw, 1
    if name == "edge_block_p0_seed32":  # a dense P0 with entries inside the groups {th, w}, {x, vx}, {y, vy}, {z, vz} only
        rng = np.random.default_rng(32)
        grp = np.array([0, 0, 0, 1, 2, 3, 0, 0, 0, 1, 2, 3])
        A = _spd(rng, 12, 0.03)
        kw["P0"] = np.where(grp[:, None] == grp[None, :], A, 0.0)
        kw["x0"] = START + 0.02 * rng.standard_normal(12)
        return make_stream(32, 300), kw, 1
    if name == "next_mpc_cov_seed5":  # predict_mpc covariance model with supplied f
        s = make_stream(5, 400)
        s["body_ref"] = s["truth"].copy()
        kw["model"] = "mpc_cov"
        kw["Q"], kw["R"] = q_r_pkl()
        return s, kw, 1
    raise KeyError(name)


ALL_CASES = [
    "cfg1_default_seed0", "default_seed11_10k", "stress_qrpkl_seed3_10k", "edge_zero_attitude_spin",
    "edge_yaw_quarter_turn", "edge_dense_noise", "edge_nonsymmetric_p0", "edge_contact_patterns",
    "edge_large_angles", "next_mpc_cov_seed5", "edge_diag_p0_seed31", "edge_block_p0_seed32",
]


def load_golden(name: str):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def stack_stream(stream):
    """[T, C] arrays -> device layout [T, C, 1]."""
    return {k: np.ascontiguousarray(v[:, :, None]) for k, v in stream.items()}
