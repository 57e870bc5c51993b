"""GPU tests of the next row after the filter and of BASELINE config 5: filter output -> feature
rows -> min-max -> windows -> GRU consumer, batched-KF arm vs reference-KF arm."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, gru_pipeline
from optistate_b200 import kf_batch
from optistate_b200.features import assemble_features, min_max, normalized_windows
from optistate_b200.synth import make_streams

pytestmark = pytest.mark.gpu


def numpy_rows(x_steps, p_world, st, idx):
    """[N, T, 60] with the oracle-side restatement, one trajectory at a time."""
    return np.stack([gru_pipeline.feature_rows(x_steps[:, :, i], st["imu_acc"][:, :, s], st["f"][:, :, s], p_world[:, :, i],
                                                st["dp"][:, :, s], st["imu"][:, :, s]) for i, s in enumerate(idx)])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_feature_rows_minmax_windows_are_exact_data_movement(dtype):
    S, T, N = 6, 77, 45  # ragged against the 32-trajectory tiles
    st = make_streams(range(60, 60 + S), T)
    res = kf_batch(st["imu"], st["p"], st["dp"], st["contact"], st["f"], n_traj=N, dtype=dtype, stream_offset=2,
                   outputs=("x_steps", "p_world_steps"))
    rows = assemble_features(res.x_steps, res.p_world_steps, st["imu"], st["f"], st["dp"], st["imu_acc"], stream_offset=2)
    idx = (np.arange(N) + 2) % S
    npdt = np.float64 if dtype == torch.float64 else np.float32
    stc = {k: v.astype(npdt) for k, v in st.items()}
    want = numpy_rows(res.x_steps.cpu().numpy(), res.p_world_steps.cpu().numpy(), stc, idx)
    assert rows.shape == (N, T, 60) and np.array_equal(rows.cpu().numpy(), want)
    flat = rows.reshape(-1, 60)
    lo, hi = min_max(flat)
    assert np.array_equal(lo.cpu().numpy(), want.reshape(-1, 60).min(axis=0)) and np.array_equal(hi.cpu().numpy(), want.reshape(-1, 60).max(axis=0))
    latent = torch.from_numpy(gru_pipeline.seeded_latent(N * T)).cuda()
    # windows over the concatenated rows (the reference's behaviour) and per trajectory
    w1 = normalized_windows(flat, lo, hi, latent, n_groups=1, seq_len=10)
    norm = ((want.reshape(-1, 60) - want.reshape(-1, 60).min(axis=0)) / (want.reshape(-1, 60).max(axis=0) - want.reshape(-1, 60).min(axis=0)))
    ref1 = np.stack([np.concatenate([norm, latent.cpu().numpy().astype(npdt)], axis=1)[i:i + 10] for i in range(N * T - 9)]).astype(np.float32)
    assert w1.shape == (1, N * T - 9, 10, 188) and np.array_equal(w1[0].cpu().numpy(), ref1, equal_nan=True)
    assert np.isnan(ref1[0, 0, 18]) and not np.isnan(ref1[..., 20]).any()  # constant columns (f_x = 0) give 0/0 exactly as in NumPy
    wn = normalized_windows(flat, lo, hi, latent, n_groups=N, seq_len=10)
    assert wn.shape == (N, T - 9, 10, 188)
    assert np.array_equal(wn[3, 5].cpu().numpy(), ref1[3 * T + 5], equal_nan=True)
    assert np.array_equal(wn[N - 1, T - 10].cpu().numpy(), ref1[(N - 1) * T + T - 10], equal_nan=True)
    w0 = normalized_windows(flat, lo, hi, None, n_groups=N, seq_len=4)
    assert w0.shape == (N, T - 3, 4, 60) and np.array_equal(w0[1, 0].cpu().numpy(), ref1[T][:4, :60], equal_nan=True)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 2e-6), (torch.float32, 2e-4)])
def test_config5_gru_rmse_parity_batched_kf_vs_reference_kf(dtype, tol):
    """BASELINE config 5: per-state RMSE of the GRU consumer fed by the batched CUDA filter vs fed by the reference
    filter (oracle port), same seeded weights and latents.  The GRU itself runs in float32 in both arms."""
    S, T = 4, 400
    st = make_streams(range(80, 80 + S), T)
    # the synthetic forces have f_x = f_y = 0 (constant columns would normalise to 0/0): give them a small lateral component
    st["f"][:, [0, 1, 3, 4, 6, 7, 9, 10], :] = 0.5 * np.random.default_rng(3).standard_normal((T, 8, S))
    ref = c_oracle.run(st, want=("x_steps", "p_world_steps"))
    rows_ref = numpy_rows(ref["x_steps"], ref["p_world_steps"], st, np.arange(S)).reshape(-1, 60)
    norm_ref, _, _ = gru_pipeline.min_max_normalise(rows_ref)
    latent = gru_pipeline.seeded_latent(S * T)
    win_ref = torch.from_numpy(gru_pipeline.windows(norm_ref, latent)).cuda()
    model = gru_pipeline.seeded_consumer().cuda()
    truth_rows = np.concatenate([st["truth"][:, :, s] for s in range(S)])

    res = kf_batch(st["imu"], st["p"], st["dp"], st["contact"], st["f"], dtype=dtype, outputs=("x_steps", "p_world_steps"))
    rows = assemble_features(res.x_steps, res.p_world_steps, st["imu"], st["f"], st["dp"], st["imu_acc"]).reshape(-1, 60)
    lo, hi = min_max(rows)
    win = normalized_windows(rows, lo, hi, torch.from_numpy(latent).cuda(), n_groups=1, seq_len=10)[0]
    assert win.shape == win_ref.shape
    assert float((win - win_ref).abs().max()) < (1e-6 if dtype == torch.float64 else 1e-3)
    with torch.no_grad():
        out_ref = model(win_ref).cpu().numpy().astype(np.float64)
        out = model(win).cpu().numpy().astype(np.float64)
    rmse_ref = gru_pipeline.per_state_rmse(out_ref, truth_rows)
    rmse = gru_pipeline.per_state_rmse(out, truth_rows)
    print("config 5 per-state RMSE (reference-KF arm):", np.array2string(rmse_ref, precision=5))
    print("config 5 |delta RMSE| max:", np.abs(rmse - rmse_ref).max())
    assert np.abs(rmse - rmse_ref).max() < tol * np.abs(rmse_ref).max()


def test_example_driver_pipeline_runs_and_is_consistent():
    """examples/driver_pipeline.py: identification -> filter -> feature rows -> windows, on the device end to end."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "driver_pipeline.py")
    spec = importlib.util.spec_from_file_location("driver_pipeline", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res, rows, windows = mod.run(n_recordings=8, n_steps=120, seq_len=10)
    assert res.algo == "sequential" and int(res.status.max()) == 0
    assert rows.shape == (8, 120, 60) and windows.shape == (8, 111, 10, 60) and windows.dtype == torch.float32
    assert torch.equal(rows[:, :, 0:12], res.x_steps.permute(2, 0, 1))  # the first 12 feature columns are the estimates
    w = windows[torch.isfinite(windows)]
    assert w.min() >= 0.0 and w.max() <= 1.0
    # the shipped driver's own call (estimate_state_mpc: the force MPC solved from the current estimate at every step)
    cl, rows_cl, windows_cl = mod.run(n_recordings=8, n_steps=40, seq_len=10, closed_loop=True)
    assert int(cl.status.max()) == 0 and not (cl.mpc_status & 7).any()
    assert rows_cl.shape == (8, 40, 60) and windows_cl.shape == (8, 31, 10, 60)
    assert torch.equal(rows_cl[:, :, 0:12], cl.x_steps.permute(2, 0, 1)) and bool(torch.isfinite(rows_cl).all())
    assert float(rows_cl[:, :, 18:30].abs().max()) > 1.0   # columns 18..29 hold the forces the MPC found
