"""Synthetic trajectory generator shaped like OptiState's real .mat recordings.

The Drive recordings the reference's raw-to-Kalman converter reads
(/root/reference/data_collection/data_conversion_raw_to_Kalman.py:46-57,443-447) are not
available offline, so every benchmark and parity test runs on streams produced here.  The
recipe is SURVEY.md section 8(d): dt = 0.01 s, trot gait with diagonal leg pairs alternating
every 25 steps, body-frame feet around the nominal stance, IMU attitude = slow sine + noise.

One *base stream* is the per-step input of one trajectory:
    imu     [T, 6]   thx thy thz dthx dthy dthz   (kalman_filter.py:108-117 uses imu[0:6])
    p       [T, 12]  body-frame foot positions, leg-major xyz
    dp      [T, 12]  body-frame foot velocities
    contact [T, 4]   0/1 stance flags
    f       [T, 12]  world-frame ground-reaction forces, leg-major xyz
    truth   [T, 12]  label stream for error summaries (noise-free attitude + nominal height)
"""
from __future__ import annotations

import numpy as np

DT = 0.01
ROBOT_MASS = 8.8
GRAVITY = 9.81
NOMINAL_FEET = np.array(
    [0.2, 0.15, -0.28, 0.2, -0.15, -0.28, -0.2, 0.15, -0.28, -0.2, -0.15, -0.28], dtype=np.float64
)
GAIT_HALF_PERIOD = 25

# Check values for seed 0; asserted by tests/test_synth.py.
SEED0_CHECK = {
    "p[0,0]": 0.2025146044218679,
    "imu[0,0]": 0.0100396157584217,
    "dp[0,0]": -0.06538286094183395,
    "f[0,2]": 42.87661229219134,
}


def make_stream(seed: int, n_steps: int) -> dict[str, np.ndarray]:
    """One base stream; draw order is part of the contract (see module docstring)."""
    rng = np.random.default_rng(seed)
    t_idx = np.arange(n_steps)
    t = t_idx * DT

    p = NOMINAL_FEET[None, :] + 0.02 * rng.standard_normal((n_steps, 12))
    dp = 0.1 * rng.standard_normal((n_steps, 12))
    imu = np.empty((n_steps, 6))
    imu[:, 0:3] = (0.05 * np.sin(2.0 * np.pi * 0.5 * t))[:, None] + 0.01 * rng.standard_normal((n_steps, 3))
    imu[:, 3:6] = 0.1 * rng.standard_normal((n_steps, 3))

    phase = (t_idx // GAIT_HALF_PERIOD) % 2
    contact = np.zeros((n_steps, 4), dtype=np.float64)
    contact[:, 0] = contact[:, 3] = (phase == 0)
    contact[:, 1] = contact[:, 2] = (phase == 1)

    f = np.zeros((n_steps, 12))
    for leg in range(4):
        f[:, 3 * leg + 2] = contact[:, leg] * (ROBOT_MASS * GRAVITY / 2.0 + rng.standard_normal(n_steps))

    truth = np.zeros((n_steps, 12))
    truth[:, 0:3] = (0.05 * np.sin(np.pi * t))[:, None]
    truth[:, 5] = 0.28
    # accelerometer channels of the recordings (imu_list[i][6:12], only copied into the GRU feature rows by the driver,
    # data_conversion_Kalman_to_Training.py:246-247); drawn from a separate generator so the filter inputs above are unchanged
    rng_acc = np.random.default_rng([seed, 1])
    imu_acc = np.array([0.0, 0.0, GRAVITY, 0.0, 0.0, 0.0])[None, :] + 0.2 * rng_acc.standard_normal((n_steps, 6))
    return {"imu": imu, "p": p, "dp": dp, "contact": contact, "f": f, "truth": truth, "imu_acc": imu_acc}


def make_streams(seeds, n_steps: int) -> dict[str, np.ndarray]:
    """Stack base streams in the device layout [T, C, S] (stream index fastest-varying)."""
    per = [make_stream(int(s), n_steps) for s in seeds]
    return {k: np.ascontiguousarray(np.stack([d[k] for d in per], axis=-1)) for k in per[0]}


def monte_carlo_noise(member_ids, q_diag: np.ndarray, r_diag: np.ndarray, nominal_every: int | None = None):
    """Per-member diagonal Q/R for the Monte-Carlo noise sweep (BASELINE.json configs 3/4).

    Member i scales every diagonal entry by 10**u, u ~ U(-0.5, 0.5), from default_rng(10**6 + i).
    Members with i < nominal_every (the first pass over the streams) keep u = v = 0 when
    nominal_every is given, so each stream has a nominal member to compare against.
    Returns (Q [12, N], R [10, N]) float64.
    """
    member_ids = np.asarray(member_ids, dtype=np.int64)
    n = member_ids.shape[0]
    q = np.empty((12, n))
    r = np.empty((10, n))
    for j, i in enumerate(member_ids):
        if nominal_every is not None and i < nominal_every:
            q[:, j] = q_diag
            r[:, j] = r_diag
            continue
        rng = np.random.default_rng(10**6 + int(i))
        q[:, j] = q_diag * 10.0 ** rng.uniform(-0.5, 0.5, 12)
        r[:, j] = r_diag * 10.0 ** rng.uniform(-0.5, 0.5, 10)
    return q, r


NOMINAL_FEET = np.array([0.2, 0.15, -0.28, 0.2, -0.15, -0.28, -0.2, 0.15, -0.28, -0.2, -0.15, -0.28])


def make_mpc_problems(n: int, seed: int = 0):
    """n force-MPC problems in device layout: x [12, n], body_ref [5, 12, n], p [12, n], contact [4, n].
    Contact patterns cycle through all 16 combinations; every fourth group of 16 demands a forward velocity that saturates
    friction and the normal-force cap."""
    rng = np.random.default_rng(seed)
    x = np.array([0.02, -0.03, 0.1, 0, 0, 0.27, 0.1, -0.1, 0.05, 0.2, -0.1, 0.0])[:, None] + 0.01 * rng.standard_normal((12, n))
    ref = np.zeros((5, 12, n))
    ref[:, 5] = 0.28
    ref[:, 0:3] = 0.02 * rng.standard_normal((5, 3, n))
    lateral = np.array([0.0, 0.5, 3.0, -2.0])[(np.arange(n) // 16) % 4]
    ref[:, 9] = lateral
    ref[:, 3] = x[3] + lateral * 0.01 * np.arange(1, 6)[:, None]
    p = NOMINAL_FEET[:, None] + 0.02 * rng.standard_normal((12, n))
    contact = ((np.arange(n)[None, :] >> np.arange(4)[:, None]) & 1).astype(np.float64)
    return x, ref, p, contact
