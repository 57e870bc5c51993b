"""The reference's OWN conversion driver, unmodified, run against the drop-in class.

/root/reference/data_collection/data_conversion_Kalman_to_Training.py (the script that steps Kalman_Filter.estimate_state_mpc over
every recorded trajectory and writes the GRU training rows) is executed byte for byte - the copy oracle/make_ref.py
stages - in a scratch tree where `kalman_filter.kalman_filter` resolves to optistate_b200.kalman_filter, `settings` to
optistate_b200.settings, matplotlib to a stub, and saved_trajectories.pkl / Q_R.pkl hold synthetic recordings.  Its output file
rnn_data.pkl must equal what the batched device path (estimate_state_mpc_batch + assemble_features) produces from the same
inputs: "the conversion driver is unchanged" as a test instead of a sentence."""
import hashlib
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import make_ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = "data_collection/data_conversion_Kalman_to_Training.py"


def _scratch_tree(tmp, trajectories, Q, R):
    ref = make_ref.root()
    src = os.path.join(ref, DRIVER)
    os.makedirs(os.path.join(tmp, "data_collection", "trajectories"))
    os.makedirs(os.path.join(tmp, "data_results"))
    os.makedirs(os.path.join(tmp, "kalman_filter"))
    os.makedirs(os.path.join(tmp, "matplotlib"))
    with open(src, "rb") as f:
        code = f.read()
    with open(os.path.join(tmp, DRIVER), "wb") as f:
        f.write(code)                                            # the driver, byte for byte
    open(os.path.join(tmp, "kalman_filter", "__init__.py"), "w").close()
    with open(os.path.join(tmp, "kalman_filter", "kalman_filter.py"), "w") as f:
        f.write("from optistate_b200.kalman_filter import Kalman_Filter  # noqa: F401\n")
    with open(os.path.join(tmp, "settings.py"), "w") as f:
        f.write("from optistate_b200.settings import INITIAL_PARAMS  # noqa: F401\n")
    open(os.path.join(tmp, "matplotlib", "__init__.py"), "w").close()
    with open(os.path.join(tmp, "matplotlib", "pyplot.py"), "w") as f:  # the driver plots every trajectory; nothing to show here
        f.write("def __getattr__(name):\n    return lambda *a, **k: None\n")
    with open(os.path.join(tmp, "data_collection", "trajectories", "saved_trajectories.pkl"), "wb") as f:
        pickle.dump(trajectories, f)
    with open(os.path.join(tmp, "data_collection", "trajectories", "Q_R.pkl"), "wb") as f:
        pickle.dump((Q, R), f)
    return hashlib.sha256(code).hexdigest()


@pytest.mark.skipif(make_ref.root() is None, reason="the reference files are not staged (oracle/_ref)")
def test_unmodified_reference_driver_runs_on_the_drop_in_and_matches_the_batched_path(tmp_path):
    from optistate_b200.features import assemble_features
    from optistate_b200.mpc import estimate_state_mpc_batch
    from optistate_b200.settings import INITIAL_PARAMS
    from optistate_b200.synth import make_streams

    n_traj, T = 2, 12
    st = make_streams(range(40, 40 + n_traj), T)
    rng = np.random.default_rng(3)
    ref_states = np.zeros((T, 12, n_traj))
    ref_states[:, 5] = 0.28
    ref_states[:, 0:3] = 0.02 * rng.standard_normal((T, 3, n_traj))
    x_start = st["truth"][0].copy()                                  # mocap_list[0]: the driver starts every filter there
    trajectories = {}
    for k in range(n_traj):
        col = lambda a, n: [a[t, :, k].reshape(n, 1).copy() for t in range(T)]  # noqa: E731
        imu12 = np.concatenate([st["imu"][:, :, k], st["imu_acc"][:, :, k]], axis=1)
        trajectories[k + 1] = {
            "p_list_est": col(st["p"], 12), "p_list_ref": col(st["p"], 12), "dp_list": col(st["dp"], 12),
            "imu_list": [imu12[t].reshape(12, 1).copy() for t in range(T)], "contact_list": col(st["contact"], 4),
            "t265_list": col(st["truth"], 12), "mocap_list": col(st["truth"], 12), "time_list": [0.01 * t for t in range(T)],
            "ref_list": [ref_states[t, :, k].reshape(12, 1).copy() for t in range(T)],
        }
    Q = np.diag([0.02, 0.01, 0.03, 0.01, 0.0002, 0.01, 0.02, 0.01, 0.01, 0.03, 0.01, 0.0001])
    R = np.diag(np.full(10, 0.02))
    digest = _scratch_tree(str(tmp_path), trajectories, Q.copy(), R.copy())
    staged = os.path.join(make_ref.root(), DRIVER)
    assert digest == hashlib.sha256(open(staged, "rb").read()).hexdigest()
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), ROOT]))
    run = subprocess.run([sys.executable, os.path.join(str(tmp_path), DRIVER)], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-3000:]
    with open(os.path.join(str(tmp_path), "data_collection", "trajectories", "rnn_data.pkl"), "rb") as f:
        out = pickle.load(f)
    assert sorted(out) == [1, 2] and os.path.isfile(os.path.join(str(tmp_path), "data_results", "p_trace_data.mat"))

    # the same conversion on the device for both trajectories at once.  What the driver sets up: x0 = mocap_list[0] (it writes
    # through the STARTING_STATE alias, so the second trajectory starts from ITS first label), Q, R with R[0:3] = 1e-4, P0 = Q;
    # body_ref = ref_list[i] held over the horizon (the class repeats a (12, 1) reference)
    R_used = R.copy()
    R_used[0, 0] = R_used[1, 1] = R_used[2, 2] = 0.0001
    body_ref = np.repeat(ref_states[:, None, :, :], 5, axis=1)
    xs, fs, mst, fst, pws = estimate_state_mpc_batch(st["imu"], st["p"], st["dp"], st["contact"], body_ref, x0=x_start, P0=Q, Q=Q, R=R_used,
                                                     return_p_world=True)
    rows = assemble_features(xs, pws, torch.from_numpy(st["imu"]), fs, torch.from_numpy(st["dp"]), torch.from_numpy(st["imu_acc"]))
    torch.cuda.synchronize()
    assert int(fst.max()) == 0 and not (mst & 7).any()
    rows = rows.cpu().numpy()
    for k in range(n_traj):
        got = np.array(out[k + 1]["state_INPUT"])
        assert got.shape == (T, 60)
        scale = np.maximum(np.abs(rows[k]).max(axis=0), 1.0)
        err = (np.abs(got - rows[k]) / scale).max()
        assert err < 1e-6, (k, err, np.unravel_index(np.argmax(np.abs(got - rows[k]) / scale), got.shape))
        assert np.array_equal(np.array(out[k + 1]["state_MOCAP"]), st["truth"][:, :, k])
    assert np.array_equal(INITIAL_PARAMS.Q, np.diag(np.diag(INITIAL_PARAMS.Q)))  # this process' settings were not touched (the driver ran in its own)
