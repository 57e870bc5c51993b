"""GPU tests THROUGH THE ctypes STUB of INTEGRATION.md: liboptistate_kf.so is loaded with ctypes.CDLL, the descriptors are the
ctypes mirrors a reference-side binding declares (tests/test_abi_cpu.py:_desc_class), and the entry points are called with raw
device pointers and a raw cudaStream_t - no PyTorch extension in the call path (torch only owns the device memory).  The results
are held to the C oracle like every other parity test."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import c_oracle, cases
from optistate_b200.synth import make_streams, monte_carlo_noise
from tests import parity
from tests.test_abi_cpu import _desc_class

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from optistate_b200 import _build

    path, _ = _build.build_all()
    so = ctypes.CDLL(path)
    so.optistate_kf_strerror.restype = ctypes.c_char_p
    return so


def _dev(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype).contiguous()


def _base_desc(lib, n_traj, n_steps, n_streams):
    Desc = _desc_class(lib)
    d = Desc()
    d.struct_size, d.abi_version = ctypes.sizeof(Desc), lib.optistate_kf_abi_version()
    d.dtype, d.algo, d.cov_model, d.phases = 0, 0, 0, 7
    d.n_traj, d.n_steps, d.n_streams = n_traj, n_steps, n_streams
    d.dt, d.mass, d.gravity = 0.01, 8.8, -9.81  # settings.py:5,11,20-23 (the oracle's defaults)
    d.inertia = (ctypes.c_double * 3)(55303643.08 / 10**9, 60119440.34 / 10**9, 105304340.05 / 10**9)
    return d


def _call(lib, d):
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.optistate_kf_batch(ctypes.byref(d), stream)
    assert rc == 0, lib.optistate_kf_strerror(rc)
    torch.cuda.synchronize()


def test_sequential_streamed_call_through_ctypes_matches_the_oracle(lib):
    """optistate_kf_batch as a foreign caller drives it: diagonal per-member noise, shared streams, a workspace (so the TMA-fed
    kernel and its measurement pre-pass run), summary + final state and covariance out."""
    S, T, N = 128, 120, 384
    st = make_streams(range(40, 40 + S), T)
    q, r = monte_carlo_noise(np.arange(N), np.diag(cases.Q_DEFAULT), np.diag(cases.R_DEFAULT))
    dev = {k: _dev(st[k]) for k in ("imu", "p", "dp", "contact", "f", "truth")}
    qd, rd, x0 = _dev(q), _dev(r), _dev(cases.START.reshape(12))
    d = _base_desc(lib, N, T, S)
    for k, t in dev.items():
        setattr(d, k, t.data_ptr())
    d.x0, d.x0_per_traj = x0.data_ptr(), 0
    d.Q, d.R, d.q_kind, d.r_kind, d.p0_kind = qd.data_ptr(), rd.data_ptr(), 2, 2, 0
    out = {"x_final": torch.empty((12, N), dtype=torch.float64, device="cuda"), "P_final": torch.empty((144, N), dtype=torch.float64, device="cuda"),
           "summary": torch.empty((52, N), dtype=torch.float64, device="cuda"), "status": torch.zeros(N, dtype=torch.int32, device="cuda")}
    for k, t in out.items():
        setattr(d, k, t.data_ptr())
    assert lib.optistate_kf_resolve_algo(ctypes.byref(d)) == 2
    nbytes = ctypes.c_size_t(0)
    assert lib.optistate_kf_workspace_bytes(ctypes.byref(d), ctypes.byref(nbytes)) == 0 and nbytes.value > 0
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes.value
    before = lib.optistate_kf_launch_count()
    _call(lib, d)
    assert lib.optistate_kf_launch_count() - before >= 1
    idx = (np.arange(N) % S).astype(np.int32)
    ref = c_oracle.run(st, N, Q=q, R=r, stream_index=idx, want=("x_final", "P_final"))
    assert parity.rel_err(out["x_final"].cpu().numpy(), ref["x_final"]) < parity.FP64_TOL
    assert parity.rel_err(out["P_final"].cpu().numpy(), ref["P_final"]) < parity.FP64_TOL
    sm = out["summary"].cpu().numpy()
    assert parity.rel_err(sm[0:12], ref["x_final"]) < parity.FP64_TOL
    assert parity.rel_err(sm[12:24], ref["P_final"].reshape(12, 12, N)[np.arange(12), np.arange(12)]) < parity.FP64_TOL
    assert int(out["status"].max()) == 0
    # the same descriptor without a workspace takes the direct-load kernel: same answer
    d.workspace, d.workspace_bytes = None, 0
    first = out["x_final"].clone()
    _call(lib, d)
    assert parity.rel_err(out["x_final"].cpu().numpy(), first.cpu().numpy()) < 1e-13


def test_joint_single_step_through_ctypes_matches_the_oracle(lib):
    """One trajectory, dense Q / R / P0 and the gain matrix out: the call the drop-in class makes per step."""
    T = 25
    st = make_streams([7], T)
    rng = np.random.default_rng(3)
    a = rng.standard_normal((12, 12)) * 0.02
    qm = np.diag(np.diag(cases.Q_DEFAULT)) + a @ a.T
    b = rng.standard_normal((10, 10)) * 0.02
    rm = np.diag(np.diag(cases.R_DEFAULT)) + b @ b.T
    dev = {k: _dev(st[k]) for k in ("imu", "p", "dp", "contact", "f")}
    qd, rd, x0 = _dev(qm.reshape(144)), _dev(rm.reshape(100)), _dev(cases.START.reshape(12))
    d = _base_desc(lib, 1, T, 1)
    for k, t in dev.items():
        setattr(d, k, t.data_ptr())
    d.x0, d.x0_per_traj = x0.data_ptr(), 0
    d.Q, d.R, d.P0, d.q_kind, d.r_kind, d.p0_kind = qd.data_ptr(), rd.data_ptr(), qd.data_ptr(), 3, 3, 3
    out = {"x_steps": torch.empty((T, 12, 1), dtype=torch.float64, device="cuda"), "P_final": torch.empty((144, 1), dtype=torch.float64, device="cuda"),
           "K_final": torch.empty((120, 1), dtype=torch.float64, device="cuda"), "status": torch.zeros(1, dtype=torch.int32, device="cuda")}
    for k, t in out.items():
        setattr(d, k, t.data_ptr())
    assert lib.optistate_kf_resolve_algo(ctypes.byref(d)) == 1
    _call(lib, d)
    ref = c_oracle.run(st, Q=qm, R=rm, P0=qm, want=("x_steps", "P_final", "K_final"))
    assert parity.state_err(out["x_steps"][:, :, 0].cpu().numpy(), ref["x_steps"][:, :, 0]) < parity.FP64_TOL
    assert parity.rel_err(out["P_final"].cpu().numpy(), ref["P_final"]) < parity.FP64_TOL
    assert parity.rel_err(out["K_final"].cpu().numpy(), ref["K_final"]) < 1e-8
    assert int(out["status"][0]) & ~8 == 0
