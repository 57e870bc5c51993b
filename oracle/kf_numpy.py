"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the reference filter step (small cases).

Follows, in execution order (all paths relative to /root/reference):
    kalman_filter/kalman_filter.py:79-105   get_odom           -> form_measurement
    kalman_filter/kalman_filter.py:108-117  set_measurements   -> form_measurement
    kalman_filter/kalman_filter.py:184-193  rotation_matrix_body_world -> rot_zyx
    misc/force_controller.py:269-291        next_state         -> propagate_mean
    kalman_filter/kalman_filter.py:119-138  predict            -> propagate_cov (model="predict")
    kalman_filter/kalman_filter.py:153-158  predict_mpc cov    -> propagate_cov (model="mpc")
    kalman_filter/kalman_filter.py:164-174  update             -> joint_update
Pinned against tests/golden/*.npz (outputs of the unmodified reference, see gen_golden.py).
"""
from __future__ import annotations

import numpy as np

SEL = np.array([0, 1, 2, 5, 6, 7, 8, 9, 10, 11])  # rows of H, kalman_filter.py:15-24
DT = 0.01  # settings.py:5
MASS = 8.8  # settings.py:11
INERTIA = np.array([55303643.08 / 10**9, 60119440.34 / 10**9, 105304340.05 / 10**9])  # settings.py:20-22
GRAVITY = -9.81  # kalman_filter.py:56
START = np.array([0, 0, 0, 0, 0, 0.28, 0, 0, 0, 0, 0, 0], dtype=np.float64)  # settings.py:25
Q_DEFAULT = np.array([0.01, 0.01, 0.01, 0.01, 0.0001, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.0001])  # settings.py:28
R_DEFAULT = np.full(10, 0.01)  # settings.py:30


def rot_zyx(a, b, c):
    """Rz(c) @ Ry(b) @ Rx(a), multiplied out (kalman_filter.py:184-193)."""
    sa, ca, sb, cb, sc, cc = np.sin(a), np.cos(a), np.sin(b), np.cos(b), np.sin(c), np.cos(c)
    return np.array(
        [
            [cc * cb, cc * (sb * sa) - sc * ca, cc * (sb * ca) + sc * sa],
            [sc * cb, sc * (sb * sa) + cc * ca, sc * (sb * ca) - cc * sa],
            [-sb, cb * sa, cb * ca],
        ]
    )


def form_measurement(imu, p, dp, contact):
    """z (10,) from one step of raw inputs; raises ValueError like the reference when no leg is in stance."""
    nc = float(np.sum(contact))
    acc = np.zeros(4)  # vx, vy, vz (swing legs!), z
    for leg in range(4):
        if contact[leg] == 1:
            acc[0] += dp[3 * leg]
            acc[1] += dp[3 * leg + 1]
            acc[3] += p[3 * leg + 2]
        if contact[leg] == 0:
            acc[2] += dp[3 * leg + 2]
    if nc == 0:
        raise ValueError("all four feet in swing: the reference builds a ragged array here (kalman_filter.py:97-103)")
    acc = -1 * acc / nc
    v_world = rot_zyx(imu[0], imu[1], imu[2]) @ acc[0:3]
    return np.array([imu[0], imu[1], imu[2], acc[3], imu[3], imu[4], imu[5], v_world[0], v_world[1], v_world[2]])


def propagate_mean(x, p, f, dt=DT, mass=MASS, inertia=INERTIA, gravity=GRAVITY):
    """next_state: returns (x', p_world, R).  The int64 `A` of force_controller.py:248-251,271 makes the
    attitude row use trunc(R^T)."""
    R = rot_zyx(x[0], x[1], x[2])
    RT_int = np.trunc(R.T)
    I_hat_inv = np.linalg.inv(R @ np.diag(inertia) @ R.T)
    p_world = np.concatenate([R @ p[3 * l : 3 * l + 3] for l in range(4)])
    xn = np.array(x, dtype=np.float64)
    xn[0:3] = x[0:3] + dt * (RT_int @ x[6:9])
    xn[3:6] = x[3:6] + dt * x[9:12]
    torque_rows = np.zeros(3)
    force_sum = np.zeros(3)
    for l in range(4):
        pw = p_world[3 * l : 3 * l + 3]
        skew = np.array([[0, -pw[2], pw[1]], [pw[2], 0, -pw[0]], [-pw[1], pw[0], 0]])
        torque_rows = torque_rows + (dt * (I_hat_inv @ skew)) @ f[3 * l : 3 * l + 3]
        force_sum = force_sum + (dt / mass) * f[3 * l : 3 * l + 3]
    xn[6:9] = x[6:9] + torque_rows
    xn[9:12] = x[9:12] + force_sum
    xn[11] += dt * gravity
    return xn, p_world, R


def propagate_cov(P, R, Q, dt=DT, model="predict"):
    """P <- F_d P F_d^T + Q with F_d = I + dt*F (predict) or exp(dt*F) element-wise (predict_mpc)."""
    F = np.zeros((12, 12))
    F[3:6, 9:12] = np.eye(3)
    F[0:3, 6:9] = R.T
    F_d = np.eye(12) + dt * F if model == "predict" else np.exp(dt * F)
    return F_d @ P @ F_d.T + Q


def joint_update(x, P, z, Rn):
    """update(): K = (P H^T) inv(S), P <- (I - K H) P, no symmetrisation (kalman_filter.py:164-174)."""
    y = z - x[SEL]
    S = P[np.ix_(SEL, SEL)] + Rn
    K = P[:, SEL] @ np.linalg.inv(S)
    x_new = x + K @ y
    KH = np.zeros((12, 12))
    KH[:, SEL] = K
    P_new = (np.eye(12) - KH) @ P
    nis = float(y @ np.linalg.solve(S, y))
    return x_new, P_new, K, float(np.trace(P_new)), float(np.trace(K)), nis


def run(stream, x0=None, P0=None, Q=None, R=None, p_checkpoint_every=0, model="predict"):
    """Same signature and outputs as ref_shim.run_reference, computed by the restatement."""
    Q = np.diag(Q_DEFAULT) if Q is None else np.asarray(Q, float)
    Rn = np.diag(R_DEFAULT) if R is None else np.asarray(R, float)
    x = START.copy() if x0 is None else np.asarray(x0, float).reshape(12).copy()
    P = Q.copy() if P0 is None else np.asarray(P0, float).copy()
    T = stream["imu"].shape[0]
    out = {k: np.empty((T, n)) for k, n in (("x", 12), ("x_model", 12), ("z", 10), ("p_world", 12))}
    out.update(p_trace=np.empty(T), k_gain=np.empty(T), nis=np.empty(T), P_ckpt={})
    for t in range(T):
        z = form_measurement(stream["imu"][t], stream["p"][t], stream["dp"][t], stream["contact"][t])
        x_pred, p_world, Rm = propagate_mean(x, stream["p"][t], stream["f"][t])
        if model == "mpc":
            br = stream["body_ref"][t]
            Rm = rot_zyx(br[0], br[1], br[2])
        P = propagate_cov(P, Rm, Q, model=model)
        x, P, K, ptr, kg, nis = joint_update(x_pred, P, z, Rn)
        out["x"][t], out["x_model"][t], out["z"][t], out["p_world"][t] = x, x_pred, z, p_world
        out["p_trace"][t], out["k_gain"][t], out["nis"][t] = ptr, kg, nis
        if p_checkpoint_every and (t + 1) % p_checkpoint_every == 0:
            out["P_ckpt"][t + 1] = P.copy()
    out["P_final"] = P
    out["K_last"] = K
    return out
