"""GPU test of the step before the filter: batched Q / R identification vs the NumPy restatement of
data_conversion_Kalman_to_Training.py:31-109 (which tests/test_host_cpu.py pins to the unmodified reference class)."""
import numpy as np
import pytest
import torch

from oracle import identify_numpy
from optistate_b200.identify import identify_noise
from optistate_b200.synth import make_streams

pytestmark = pytest.mark.gpu


def ground_truth(st, seed):
    rng = np.random.default_rng(seed)
    gt = st["truth"] + 0.02 * rng.standard_normal(st["truth"].shape)
    gt[:, 6:12] = 0.1 * rng.standard_normal(gt[:, 6:12].shape)
    return gt


@pytest.mark.parametrize("alias", [False, True])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-4)])
def test_identify_noise_matches_restatement(dtype, tol, alias):
    S, T = 9, 257
    st = make_streams(range(900, 900 + S), T)
    gt = ground_truth(st, 1)
    q, r, status = identify_noise(gt, st["imu"], st["p"], st["dp"], st["contact"], st["f"], dtype=dtype, alias_last_measurement=alias)
    for s in range(S):
        q_ref, r_ref = identify_numpy.identify(gt[:, :, s], st["imu"][:, :, s], st["p"][:, :, s], st["dp"][:, :, s], st["contact"][:, :, s],
                                               st["f"][:, :, s], alias_last_measurement=alias)
        assert np.abs(q[:, s].cpu().numpy() / q_ref - 1).max() < tol
        assert np.abs(r[:, s].cpu().numpy() / r_ref - 1).max() < tol
    assert int(status.max()) == 0
    # shared streams with an explicit mapping
    idx = np.array([3, 3, 0, 8], dtype=np.int32)
    q2, r2, _ = identify_noise(gt, st["imu"], st["p"], st["dp"], st["contact"], st["f"], dtype=dtype, n_traj=4, stream_index=idx,
                               alias_last_measurement=alias)
    assert torch.equal(q2[:, 0], q2[:, 1]) and torch.equal(q2[:, 2], q[:, 0]) and torch.equal(r2[:, 3], r[:, 8])
