"""The constants the filter reads, under the names the reference's drivers use (`settings.INITIAL_PARAMS`,
/root/reference/settings.py:2-31) and with the same aliasing: `P` is the same object as `Q` (settings.py:31), and
`Kalman_Filter().x` is the same object as `STARTING_STATE` (kalman_filter.py:10)."""
import numpy as np

_INERTIA_DIAGONAL_E9 = (55303643.08, 60119440.34, 105304340.05)  # body-frame xx, yy, zz in kg m^2 x 1e9
_Q_DIAGONAL = [1e-2] * 4 + [1e-4] + [1e-2] * 6 + [1e-4]           # thx thy thz x | y | z wx wy wz vx vy | vz
_R_DIAGONAL = [1e-2] * 10                                          # th_imu(3) z_odom w_imu(3) v_odom(3)


class INITIAL_PARAMS:
    DT = DT_mpc = 1e-2                       # filter / MPC step in seconds
    ROBOT_HEIGHT, ROBOT_MASS, GRAVITY = 0.28, 8.8, -9.81   # GRAVITY: kalman_filter.py:56
    KF_FREQUENCY = 500
    DATA_CUTOFF_START, DATA_CUTOFF_END = 430, 4494
    VISUALIZE_DATA_CONVERSION = False        # plotting is out of scope here
    Px, Py, Pz = (v / 1e9 for v in _INERTIA_DIAGONAL_E9)
    INERTIA_ROT = np.diag([Px, Py, Pz])
    STARTING_STATE = np.zeros((12, 1))
    STARTING_STATE[5, 0] = ROBOT_HEIGHT
    Q = np.diag(np.asarray(_Q_DIAGONAL, dtype=float))
    R = np.diag(np.asarray(_R_DIAGONAL, dtype=float))
    P = Q
