"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the reference's Q / R identification branch
(/root/reference/data_collection/data_conversion_Kalman_to_Training.py:31-109) for one recording, with the MPC forces
supplied.  `alias_last_measurement=True` reproduces the reference literally: `measurement_data.append(KF.z)` (:74) stores
the same array object every step, so all entries equal the last measurement."""
from __future__ import annotations

import numpy as np

from . import kf_numpy


def identify(gt, imu, p, dp, contact, f, alias_last_measurement=False):
    """All arguments [T, C].  Returns (q_diag [12], r_diag [10])."""
    T = gt.shape[0]
    model, meas = [], []
    for i in range(T - 1):
        x_model, _, _ = kf_numpy.propagate_mean(gt[i], p[i], f[i])          # KF.x = ground_truth; predict_mpc -> x_model  (:59-66)
        model.append(x_model)
        meas.append(kf_numpy.form_measurement(imu[i + 1], p[i + 1], dp[i + 1], contact[i + 1]))  # :68-72
    if alias_last_measurement:
        meas = [meas[-1]] * len(meas)
    e_x = np.array([gt[i + 1] - model[i] for i in range(T - 1)])              # :81-83
    e_z = np.array([gt[i + 1][kf_numpy.SEL] - meas[i] for i in range(T - 1)])  # :94-98
    return np.var(e_x, axis=0), np.var(e_z, axis=0)                           # :88-89, :104-106


def identify_with_reference_class(gt, imu, p, dp, contact, f):
    """The same loop driven through the UNMODIFIED reference class and next_state (build container only): cross-checks the
    restatement above including the aliasing of KF.z."""
    from . import ref_shim

    KF_cls, fc, _ = ref_shim.load()
    kf = KF_cls()
    model_data, measurement_data, ground_truth_data = [], [], []
    T = gt.shape[0]
    for i in range(T - 1):
        kf.x = gt[i].reshape(12, 1).copy()
        kf.x = fc.next_state(kf.x, p[i].reshape(12, 1).copy(), f[i].reshape(12, 1), kf.dt)
        kf.x_model = kf.x.copy()
        odom = kf.get_odom(p[i + 1].reshape(12, 1), dp[i + 1].reshape(12, 1), contact[i + 1].reshape(4, 1), imu[i + 1].reshape(6, 1))
        kf.set_measurements(imu[i + 1].reshape(6, 1), odom)
        ground_truth_data.append(gt[i + 1].reshape(12, 1))
        measurement_data.append(kf.z)        # the reference appends the same object every step
        model_data.append(kf.x_model)
    e_x = np.array([ground_truth_data[i] - model_data[i] for i in range(T - 1)]).reshape(T - 1, -1)
    sel = kf_numpy.SEL
    e_z = np.array([ground_truth_data[i][sel] - measurement_data[i] for i in range(T - 1)]).reshape(T - 1, -1)
    return np.var(e_x, axis=0), np.var(e_z, axis=0)
