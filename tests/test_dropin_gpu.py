"""GPU tests of the drop-in `Kalman_Filter` class: the reference driver's call sequence, side effects and error
conventions, checked against the goldens minted from the unmodified reference class."""
import numpy as np
import pytest

from oracle import cases
from optistate_b200 import Kalman_Filter
from optistate_b200.settings import INITIAL_PARAMS

pytestmark = pytest.mark.gpu


def drive(kf, stream, n_steps, mode="predict"):
    xs, pws, tr, kg = [], [], [], []
    for t in range(n_steps):
        imu = stream["imu"][t].reshape(6, 1).copy()
        p = stream["p"][t].reshape(12, 1).copy()
        dp = stream["dp"][t].reshape(12, 1).copy()
        contact = stream["contact"][t].reshape(4, 1).copy()
        f = stream["f"][t].reshape(12, 1).copy()
        if mode == "predict":  # the north-star path: get_odom -> set_measurements -> predict(p, f) -> update
            kf.set_measurements(imu, kf.get_odom(p, dp, contact, imu))
            kf.predict(p, f)
            kf.update()
            x = kf.x
        else:  # the shipped driver's call (data_conversion_Kalman_to_Training.py:199) with supplied forces
            x = kf.estimate_state_mpc(imu, p, dp, stream["body_ref"][t].reshape(12, 1), contact, f=f)
        xs.append(x.reshape(12).copy()); pws.append(p.reshape(12).copy()); tr.append(kf.P_trace); kg.append(kf.K_gain)
    return np.array(xs), np.array(pws), np.array(tr), np.array(kg)


def test_reference_call_sequence_matches_golden():
    stream, kw, _ = cases.build("cfg1_default_seed0")
    g = cases.load_golden("cfg1_default_seed0")
    kf = Kalman_Filter()
    start_alias = kf.x
    kf.x = kf.x.copy()
    n = 60
    xs, pws, tr, kg = drive(kf, stream, n)
    assert np.abs(xs - g["x"][:n]).max() < 1e-12
    assert np.abs(pws - g["p_world"][:n]).max() < 1e-13  # predict() rotated the caller's p in place
    assert np.abs(tr / g["p_trace"][:n] - 1).max() < 1e-12 and np.abs(kg / g["k_gain"][:n] - 1).max() < 1e-12
    assert kf.x.shape == (12, 1) and kf.P.shape == (12, 12) and kf.K.shape == (12, 10) and kf.x_model.shape == (12, 1)
    assert np.array_equal(start_alias, INITIAL_PARAMS.STARTING_STATE) and start_alias[5, 0] == 0.28  # rebinding, not mutation
    assert np.allclose(kf.F_d, np.eye(12) + kf.dt * kf.F) and np.allclose(kf.F[3:6, 9:12], np.eye(3))


def test_driver_style_setup_with_q_r_fixture_and_dense_matrices():
    """The driver assigns Q, R (dense np.diag matrices), overrides R[0..2] and sets P = deepcopy(Q)
    (data_conversion_Kalman_to_Training.py:136-144)."""
    stream, kw, _ = cases.build("stress_qrpkl_seed3_10k")
    g = cases.load_golden("stress_qrpkl_seed3_10k")
    kf = Kalman_Filter()
    kf.x = kw["x0"].reshape(12, 1).copy()
    kf.Q, kf.R = kw["Q"], kw["R"]
    kf.P = kw["Q"].copy()
    xs, _, tr, kg = drive(kf, stream, 40)
    st = g["steps"][g["steps"] < 40]
    scale = g["x_absmax"]
    assert (np.abs(xs[st] - g["x"][: len(st)]) / scale).max() < 1e-9
    assert np.abs(tr[st] / g["p_trace"][: len(st)] - 1).max() < 1e-9


def test_estimate_state_mpc_with_supplied_forces_matches_golden():
    stream, kw, _ = cases.build("next_mpc_cov_seed5")
    g = cases.load_golden("next_mpc_cov_seed5")
    kf = Kalman_Filter()
    kf.x = kw["x0"].reshape(12, 1).copy()
    kf.Q, kf.R, kf.P = kw["Q"], kw["R"], kw["Q"].copy()
    xs, pws, tr, kg = drive(kf, stream, 30, mode="mpc")
    assert (np.abs(xs - g["x"][:30]) / g["x_absmax"]).max() < 1e-9
    assert np.abs(pws - g["p_world"][:30]).max() < 1e-10  # attitude error of the ill-conditioned mpc model (~1e-12) rotates the feet
    assert np.abs(kg / g["k_gain"][:30] - 1).max() < 1e-8
    assert kf.f.shape == (12, 1)
    provider_calls = []
    kf2 = Kalman_Filter(force_provider=lambda p, body_ref, contact, x: provider_calls.append(1) or np.zeros((12, 6)))
    kf2.x = kf2.x.copy()
    kf2.estimate_state_mpc(stream["imu"][0].reshape(6, 1), stream["p"][0].reshape(12, 1).copy(), stream["dp"][0].reshape(12, 1),
                           stream["body_ref"][0].reshape(12, 1), stream["contact"][0].reshape(4, 1))
    assert provider_calls == [1] and kf2.f.shape == (12, 6)  # the MPC horizon matrix is kept as KF.f, column 0 is applied


def test_error_conventions():
    stream, _, _ = cases.build("cfg1_default_seed0")
    kf = Kalman_Filter()
    kf.x = kf.x.copy()
    imu, p, dp = stream["imu"][0].reshape(6, 1), stream["p"][0].reshape(12, 1), stream["dp"][0].reshape(12, 1)
    with pytest.raises(ValueError):  # kalman_filter.py:97-103 builds a ragged array when no foot is in stance
        kf.get_odom(p, dp, np.zeros((4, 1)), imu)
    odom = kf.get_odom(p, dp, np.array([1, 0, 0, 1.0]).reshape(4, 1), imu)
    assert odom.shape == (4, 1)
    kf.P = np.zeros((12, 12))
    kf.R = np.zeros((10, 10))
    with pytest.raises(np.linalg.LinAlgError):  # singular S (kalman_filter.py:168)
        kf.update()
    # an INDEFINITE but regular S does not raise in the reference (np.linalg.inv only fails on an exactly singular matrix): the
    # device path takes the pivoted inverse there too and returns the reference's numbers
    kf3 = Kalman_Filter()
    kf3.x = np.arange(12, dtype=float).reshape(12, 1) * 0.01
    kf3.P = np.diag(np.linspace(0.01, 0.02, 12))
    kf3.R = np.diag([0.01, 0.01, 0.01, -1.0, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01])
    kf3.z = np.linspace(-0.1, 0.1, 10).reshape(10, 1)
    H = np.asarray(kf3.H, float)
    S = H @ kf3.P @ H.T + kf3.R
    K = kf3.P @ H.T @ np.linalg.inv(S)
    want_x = kf3.x + K @ (kf3.z - H @ kf3.x)
    want_P = (np.eye(12) - K @ H) @ kf3.P
    kf3.update()
    assert np.abs(kf3.x - want_x).max() < 1e-12 and np.abs(kf3.P - want_P).max() < 1e-13 and np.abs(kf3.K - K).max() < 1e-12
    # the attributes of the reference's MPC set-up exist with its shapes (kalman_filter.py:37-43,73-77)
    assert kf3.p_mpc.shape == (12, 6) and kf3.body_mpc.shape == (12, 6) and kf3.contact_mpc.shape == (4, 5)
    assert kf3.zero_mat.shape == (3, 3) and np.array_equal(kf3.identity, np.eye(3)) and np.allclose(kf3.identity_m, np.eye(3) / kf3.m)


@pytest.mark.parametrize("transfer", ["copy", "store", "mapped"])
def test_every_transfer_mode_gives_the_same_numbers(transfer):
    """How the bytes of a call cross the host interface (asynchronous copies, zero-copy stores, zero-copy loads and stores) is
    plumbing: every mode reproduces the golden of the unmodified reference class."""
    stream, kw, _ = cases.build("cfg1_default_seed0")
    g = cases.load_golden("cfg1_default_seed0")
    kf = Kalman_Filter()
    kf.transfer = transfer
    kf.x = kf.x.copy()
    xs, pws, tr, kg = drive(kf, stream, 25)
    assert np.abs(xs - g["x"][:25]).max() < 1e-12 and np.abs(pws - g["p_world"][:25]).max() < 1e-13
    assert np.abs(tr / g["p_trace"][:25] - 1).max() < 1e-12 and np.abs(kg / g["k_gain"][:25] - 1).max() < 1e-12


def test_update_computed_ahead_is_only_used_when_nothing_changed():
    """predict() launches the update that normally follows it in the same call; update() may hand that result out only if x, P, z
    and R are what that launch read.  A caller that touches any of them between the two calls gets an update of what it set."""
    from optistate_b200 import _native as nv

    stream, _, _ = cases.build("cfg1_default_seed0")
    cols = {k: stream[k][3].reshape(-1, 1).copy() for k in ("imu", "p", "dp", "contact", "f")}

    def fresh():
        kf = Kalman_Filter()
        kf.x = kf.x.copy()
        kf.P = kf.P.copy()
        kf.set_measurements(cols["imu"], kf.get_odom(cols["p"], cols["dp"], cols["contact"], cols["imu"]))
        kf.predict(cols["p"].copy(), cols["f"])
        return kf

    # untouched: no launch in update(), and the numbers are those of an update launched on its own
    a, b = fresh(), fresh()
    before = nv.ext().launch_count()
    a.update()
    assert nv.ext().launch_count() == before
    b._ahead = None
    b.update()
    assert nv.ext().launch_count() == before + 1
    assert np.array_equal(a.x, b.x) and np.array_equal(a.P, b.P) and np.array_equal(a.K, b.K) and a.K_gain == b.K_gain and a.P_trace == b.P_trace

    def host_update(kf):
        H = np.asarray(kf.H, float)
        x, P, z, R = (np.array(v, dtype=float) for v in (kf.x, kf.P, kf.z, kf.R))
        K = P @ H.T @ np.linalg.inv(H @ P @ H.T + R)
        return x + K @ (z - H @ x), (np.eye(12) - K @ H) @ P

    # every input of the update, changed in place or rebound after the prediction, is honoured
    for change in ("z_in_place", "x_in_place", "P_rebound", "R_rebound"):
        kf = fresh()
        if change == "z_in_place":
            kf.z[3] += 0.05
        elif change == "x_in_place":
            kf.x[9:12] += 0.1
        elif change == "P_rebound":
            kf.P = kf.P * 1.5
        else:
            kf.R = np.asarray(kf.R) * 2.0
        want_x, want_P = host_update(kf)
        before = nv.ext().launch_count()
        kf.update()
        assert nv.ext().launch_count() == before + 1, change
        assert np.abs(kf.x - want_x).max() < 1e-11 and np.abs(kf.P - want_P).max() < 1e-12, change
