"""Mirror of the reference's settings.INITIAL_PARAMS for the constants the filter reads
(/root/reference/settings.py:2-31).  Same attribute names, same aliasing: P is the same object as Q
(settings.py:31) and Kalman_Filter().x is the same object as STARTING_STATE (kalman_filter.py:10)."""
import numpy as np


class INITIAL_PARAMS:
    DT_mpc = 0.01
    DT = 0.01
    ROBOT_HEIGHT = 0.28
    ROBOT_MASS = 8.8
    KF_FREQUENCY = 500
    DATA_CUTOFF_START = 430
    DATA_CUTOFF_END = 4494
    VISUALIZE_DATA_CONVERSION = False
    Px = 55303643.08 / (10 ** 9)
    Py = 60119440.34 / (10 ** 9)
    Pz = 105304340.05 / (10 ** 9)
    INERTIA_ROT = np.array([[Px, 0, 0], [0, Py, 0], [0, 0, Pz]])
    STARTING_STATE = np.array([0.0, 0.0, 0.0, 0.0, 0.0, ROBOT_HEIGHT, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]).reshape(12, 1)
    Q = np.diag([0.01, 0.01, 0.01, 0.01, 0.0001, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.0001])
    R = np.diag([0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01])
    P = Q
    GRAVITY = -9.81  # kalman_filter.py:56
