"""Batched entry point: filters N independent trajectories over T steps in one kernel launch.

This is the new call the reference does not have; per trajectory and per step it computes exactly what the
reference driver computes with
    KF.set_measurements(imu, KF.get_odom(p, dp, contact, imu)); KF.predict(p, f); KF.update()
(/root/reference/kalman_filter/kalman_filter.py:79-138,164-174; loop shape of
/root/reference/data_collection/data_conversion_Kalman_to_Training.py:193-201).

Layouts are structure-of-arrays with the trajectory / stream index fastest-varying (coalesced on the device):
per-step inputs [T, C, S], per-step outputs [T, C, N], per-trajectory arrays [C, N].  Trajectory i reads base
stream `stream_index[i]`, or `(i + stream_offset) % S` when no index is given, so thousands of Monte-Carlo members
can share a few input streams.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from . import _native as nv
from .settings import INITIAL_PARAMS

_OUT_ALIASES = {"p_trace": "p_trace_steps", "k_gain": "k_gain_steps", "p_checkpoints": "P_ckpt", "nis": "nis_steps"}
_OUT_SHAPES = {
    "x_steps": lambda T, N, K: (T, 12, N), "x_model_steps": lambda T, N, K: (T, 12, N),
    "p_world_steps": lambda T, N, K: (T, 12, N), "z_steps": lambda T, N, K: (T, 10, N),
    "p_trace_steps": lambda T, N, K: (T, N), "k_gain_steps": lambda T, N, K: (T, N), "nis_steps": lambda T, N, K: (T, N),
    "P_ckpt": lambda T, N, K: (K, 144, N), "x_final": lambda T, N, K: (12, N), "P_final": lambda T, N, K: (144, N),
    "K_final": lambda T, N, K: (120, N), "summary": lambda T, N, K: (nv.SUMMARY_ROWS, N),
}
_ALGOS = {"auto": nv.ALGO_AUTO, "joint": nv.ALGO_JOINT, "sequential": nv.ALGO_SEQUENTIAL}
_ALGO_NAMES = {nv.ALGO_JOINT: "joint", nv.ALGO_SEQUENTIAL: "sequential"}

# decoupled groups of the predict() model (include/optistate_kf.h OPTI_KF_FLAG_*): {th, w}, {x, vx}, {y, vy}, {z, vz}
_GROUP = torch.tensor([0, 0, 0, 1, 2, 3, 0, 0, 0, 1, 2, 3])
_CROSS_GROUP = _GROUP[:, None] != _GROUP[None, :]

SUMMARY_FIELDS = {
    "x_final": slice(0, 12), "p_diag": slice(12, 24), "rmse_truth": slice(24, 36), "rms_dev_nominal": slice(36, 48),
    "mean_nis": 48, "p_trace": 49, "k_gain": 50, "max_nis_sqrt": 51,
}


@dataclass
class KfBatchResult:
    """Device tensors of one kf_batch call.  `status[i]` is a bit mask (see optistate_kf.h OPTI_KF_ST_*)."""
    algo: str
    n_traj: int
    n_steps: int
    tensors: Dict[str, torch.Tensor] = field(default_factory=dict)
    status: Optional[torch.Tensor] = None

    def __getattr__(self, name):
        t = self.__dict__.get("tensors", {})
        if name in t:
            return t[name]
        raise AttributeError(name)

    def P_matrix(self, which: str = "P_final") -> torch.Tensor:
        """[..., 144, N] -> [..., N, 12, 12]."""
        t = self.tensors[which]
        return t.movedim(-1, -2).reshape(*t.shape[:-2], t.shape[-1], 12, 12)

    def summary_field(self, name: str) -> torch.Tensor:
        return self.tensors["summary"][SUMMARY_FIELDS[name]]


def _as_device(a, dtype, device, non_blocking=True):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.device != device or t.dtype != dtype:
        t = t.to(device=device, dtype=dtype, non_blocking=non_blocking)
    return t.contiguous()


def _stream_tensor(a, C, name, dtype, device):
    t = _as_device(a, dtype, device)
    if t is None:
        return None, None, None
    if t.dim() == 2:
        t = t.unsqueeze(-1)
    if t.dim() != 3 or t.shape[1] != C:
        raise ValueError(f"{name} must have shape [T, {C}, S], got {tuple(t.shape)}")
    return t.contiguous(), t.shape[0], t.shape[2]


def _checked_stream_index(stream_index, N: int, S: int, device) -> torch.Tensor:
    """The kernels use the entries as gather offsets, so they are range-checked here (one small reduction; the gather path is not
    the streamed hot path)."""
    t = _as_device(stream_index, torch.int32, device).reshape(N)
    if N > 0:
        lo, hi = (int(v) for v in torch.aminmax(t))
        if lo < 0 or hi >= S:
            raise ValueError(f"stream_index entries must lie in [0, {S}), got [{lo}, {hi}]")
    return t


def _is_diagonal(m: torch.Tensor) -> bool:
    return bool(torch.count_nonzero(m - torch.diag(torch.diagonal(m))) == 0)


def _noise(a, n, N, name, dtype, device, kind=None):
    """Classifies a noise / covariance argument -> (tensor, kind).  Shapes: [n] shared diagonal, [n, N] diagonal per
    trajectory, [n, n] shared dense, [n*n, N] or [n, n, N] dense per trajectory.  A shared dense matrix that is
    exactly diagonal (np.diag(...), as the reference builds them) is passed on as a diagonal."""
    t = _as_device(a, dtype, device)
    shp = tuple(t.shape)
    if kind is None:
        if shp == (n,):
            kind = nv.MAT_DIAG
        elif shp == (n, n):  # also the [n, N] layout when N == n: pass kind= to disambiguate
            kind = nv.MAT_DENSE
        elif shp == (n, N):
            kind = nv.MAT_DIAG_PER
        elif shp in ((n * n, N), (n, n, N)):
            kind = nv.MAT_DENSE_PER
        else:
            raise ValueError(f"{name}: unsupported shape {shp} for n={n}, N={N}")
    if kind == nv.MAT_DENSE and _is_diagonal(t.reshape(n, n)):
        t, kind = torch.diagonal(t.reshape(n, n)).contiguous(), nv.MAT_DIAG
    return t.contiguous(), kind


def kf_batch(
    imu, p, dp, contact, f, x0=None, P0=None, Q=None, R=None, *,
    n_traj: Optional[int] = None, dtype: torch.dtype = torch.float64, stream_index=None, stream_offset: int = 0,
    outputs: Iterable[str] = ("x_steps",), ckpt_every: int = 0, truth=None, nominal=None, body_ref=None, z=None,
    algo: str = "auto", cov_model: str = "predict", q_kind=None, r_kind=None, p0_kind=None,
    dt: float = INITIAL_PARAMS.DT_mpc, mass: float = INITIAL_PARAMS.ROBOT_MASS, inertia=None,
    gravity: float = INITIAL_PARAMS.GRAVITY, device=None, out: Optional[Dict[str, torch.Tensor]] = None,
    summary_peers=None, p0_is_symmetric: Optional[bool] = None, structure: str = "auto", packed: bool = True,
) -> KfBatchResult:
    """Runs the Kalman filter over N trajectories x T steps on the current CUDA device.

    imu [T,6,S], p [T,12,S], dp [T,12,S], contact [T,4,S], f [T,12,S]: base input streams (torch CUDA tensors are
    used in place; NumPy arrays / CPU tensors are copied to the device).  `z` [T,10,S] may be given instead of
    imu/dp/contact (pre-formed measurements, see kf_measure).
    x0: [12] or [12,N] (default STARTING_STATE);  Q, R: see _noise (defaults INITIAL_PARAMS);  P0: None = Q.
    outputs: any of x_steps, x_model_steps, p_world_steps, z_steps, p_trace_steps (alias p_trace), k_gain_steps
    (k_gain), nis_steps, P_ckpt (p_checkpoints, needs ckpt_every), x_final, P_final, final (= both), K_final, summary.
    algo: "auto" | "sequential" | "joint" (see include/optistate_kf.h).  cov_model: "predict" | "mpc".
    out: optional preallocated output tensors by canonical name.
    p0_is_symmetric: skips the (synchronising) symmetry test of a dense P0 when the caller knows the answer, e.g. when P0
    is the P_final of a previous sequential call.
    structure: "auto" | "full" - the sequential kernels exploit that predict()'s F_d only couples attitude with body rate
    and each position with the velocity of its axis, so with a P0 without entries across those groups (None, diagonal, or
    a dense P0 whose cross-group entries are all zero - checked here) the cross-group entries of P are exact zeros at every
    step, in the reference too, and are neither stored nor multiplied (bit-identical results, less than half the
    arithmetic).  "full" keeps all 78 packed entries regardless (tests, comparison).
    packed: FP32 only - False keeps one trajectory per thread where the packed two-per-thread kernel (FFMA2) would be chosen.
    summary_peers: an optistate_b200.peer.PeerSummary - the summary is then written into this rank's columns of the
    job-wide [52, n_total] array on EVERY GPU of the box by the filter kernel itself (fused all-gather over NVLink peer
    stores); `result.summary` is the local column block, `summary_peers.tensor` the gathered array once
    `summary_peers.wait()` has run.
    """
    nv.require_cuda()
    ext = nv.ext()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dtype not in (torch.float64, torch.float32):
        raise ValueError("dtype must be torch.float64 or torch.float32")

    tensors: Dict[str, torch.Tensor] = {}
    T = S = None
    for name, arr, C in (("imu", imu, 6), ("p", p, 12), ("dp", dp, 12), ("contact", contact, 4), ("f", f, 12),
                         ("z_in", z, 10), ("body_ref", body_ref, 12), ("truth", truth, 12), ("nominal", nominal, 12)):
        t, tT, tS = _stream_tensor(arr, C, name, dtype, device)
        if t is None:
            continue
        if T is None:
            T, S = tT, tS
        elif (tT, tS) != (T, S):
            raise ValueError(f"{name} has [T,S]=({tT},{tS}), expected ({T},{S})")
        tensors[name] = t
    if T is None:
        raise ValueError("no input streams given")
    N = int(S if n_traj is None else n_traj)
    phases = nv.PHASE_ALL if z is None else (nv.PHASE_PREDICT | nv.PHASE_UPDATE)

    x0_t = _as_device(INITIAL_PARAMS.STARTING_STATE.reshape(12) if x0 is None else x0, dtype, device)
    x0_t = x0_t.reshape(12) if x0_t.numel() == 12 else x0_t.reshape(12, N)
    tensors["x0"] = x0_t.contiguous()
    tensors["Q"], qk = _noise(INITIAL_PARAMS.Q if Q is None else Q, 12, N, "Q", dtype, device, q_kind)
    tensors["R"], rk = _noise(INITIAL_PARAMS.R if R is None else R, 10, N, "R", dtype, device, r_kind)
    pk = nv.MAT_NONE
    p0_symmetric = True
    if structure not in ("auto", "full"):
        raise ValueError("structure must be 'auto' or 'full'")
    flags = nv.FLAG_FULL_COVARIANCE if structure == "full" else 0
    if not packed:
        flags |= nv.FLAG_SCALAR_FP32
    if P0 is not None:
        tensors["P0"], pk = _noise(P0, 12, N, "P0", dtype, device, p0_kind)
        if pk in (nv.MAT_DENSE, nv.MAT_DENSE_PER):
            m = tensors["P0"].reshape(12, 12, -1)
            p0_symmetric = bool(torch.equal(m, m.transpose(0, 1))) if p0_is_symmetric is None else bool(p0_is_symmetric)
            if structure == "auto" and p0_is_symmetric is None and p0_symmetric and cov_model != "mpc":
                # (the test synchronises, like the symmetry test; callers that pass p0_is_symmetric skip both)
                if not bool(torch.count_nonzero(m[_CROSS_GROUP.to(device)])):
                    flags |= nv.FLAG_P0_DECOUPLED
    if stream_index is not None:
        tensors["stream_index"] = _checked_stream_index(stream_index, N, S, device)

    want = []
    for o in outputs:
        o = _OUT_ALIASES.get(o, o)
        want.extend(["x_final", "P_final"] if o == "final" else [o])
    n_ckpt = T // ckpt_every if ckpt_every else 0
    for o in want:
        if o not in _OUT_SHAPES:
            raise ValueError(f"unknown output '{o}'")
        shape = _OUT_SHAPES[o](T, N, n_ckpt)
        if o == "summary" and summary_peers is not None:
            if summary_peers.dtype != dtype or summary_peers.n_local != N or summary_peers.device != device:
                raise ValueError("summary_peers was set up for another dtype, device or shard size")
            tensors[o] = summary_peers.tensor
        elif out is not None and o in out:
            if tuple(out[o].shape) != shape or out[o].dtype != dtype:
                raise ValueError(f"out['{o}'] must be {shape} {dtype}")
            tensors[o] = out[o]
        else:
            tensors[o] = torch.empty(shape, dtype=dtype, device=device)
    status = out["status"] if out is not None and "status" in out else torch.zeros(N, dtype=torch.int32, device=device)
    tensors["status"] = status

    cfg = dict(dtype=nv.F64 if dtype == torch.float64 else nv.F32, algo=_ALGOS[algo],
               cov_model=nv.COV_MPC if cov_model == "mpc" else nv.COV_PREDICT, phases=phases, n_traj=N, n_steps=T,
               n_streams=S, stream_offset=int(stream_offset), x0_per_traj=int(x0_t.dim() == 2), p0_kind=pk, q_kind=qk,
               r_kind=rk, ckpt_every=int(ckpt_every), want_K=int("K_final" in want), flags=flags)
    if summary_peers is not None and "summary" in want:
        cfg.update(summary_peers.cfg())
    if cfg["algo"] == nv.ALGO_AUTO:
        # a symmetric dense P0 is fine for the packed-symmetric kernel; a non-symmetric one needs the joint form
        probe = dict(cfg, algo=nv.ALGO_SEQUENTIAL)
        cfg["algo"] = nv.ALGO_SEQUENTIAL if (p0_symmetric and ext.kf_resolve_algo(probe) == nv.ALGO_SEQUENTIAL) else nv.ALGO_JOINT
    elif cfg["algo"] == nv.ALGO_SEQUENTIAL and not p0_symmetric:
        raise ValueError("algo='sequential' needs a symmetric P0")
    inertia = np.diag(INITIAL_PARAMS.INERTIA_ROT) if inertia is None else np.asarray(inertia, float).reshape(3)
    consts = dict(dt=float(dt), mass=float(mass), inertia0=float(inertia[0]), inertia1=float(inertia[1]),
                  inertia2=float(inertia[2]), gravity=float(gravity))
    with torch.cuda.device(device):
        # scratch for the streamed SEQUENTIAL path (measurement pre-pass + TMA-fed kernel); 0 bytes when unused
        ws_bytes = ext.kf_batch(dict(cfg, query_workspace=1), consts, tensors)
        if ws_bytes < 0:
            nv.check(int(ws_bytes), "optistate_kf_workspace_bytes")
        if ws_bytes > 0:
            ws = out.get("workspace") if out is not None else None
            if ws is None or ws.numel() < ws_bytes:
                ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=device)
            tensors["workspace"] = ws
        nv.check(int(ext.kf_batch(cfg, consts, tensors)), "optistate_kf_batch")
    res = {k: tensors[k] for k in want}
    if summary_peers is not None and "summary" in res:
        res["summary"] = summary_peers.local
    return KfBatchResult(algo=_ALGO_NAMES[cfg["algo"]], n_traj=N, n_steps=T, tensors=res, status=status)


def kf_measure(imu, p, dp, contact, *, dtype: torch.dtype = torch.float64, device=None, want_odom: bool = False):
    """Batched get_odom + set_measurements (kalman_filter.py:79-117): returns z [T,10,S] (and odom [T,4,S], status [S])."""
    nv.require_cuda()
    ext = nv.ext()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    tensors = {}
    T = S = None
    for name, arr, C in (("imu", imu, 6), ("p", p, 12), ("dp", dp, 12), ("contact", contact, 4)):
        t, tT, tS = _stream_tensor(arr, C, name, dtype, device)
        if T is None:
            T, S = tT, tS
        elif (tT, tS) != (T, S):
            raise ValueError(f"{name} has [T,S]=({tT},{tS}), expected ({T},{S})")
        tensors[name] = t
    tensors["z"] = torch.empty((T, 10, S), dtype=dtype, device=device)
    if want_odom:
        tensors["odom"] = torch.empty((T, 4, S), dtype=dtype, device=device)
    tensors["status"] = torch.zeros(S, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        nv.check(ext.kf_measure(nv.F64 if dtype == torch.float64 else nv.F32, T, S, tensors), "optistate_kf_measure")
    if want_odom:
        return tensors["z"], tensors["odom"], tensors["status"]
    return tensors["z"], tensors["status"]


def fma_peak(dtype: torch.dtype = torch.float64, fma_per_thread: int = 1 << 16):
    """Measured FMA issue peak of the current device: (FLOP/s, seconds of the timed launch)."""
    nv.require_cuda()
    return nv.ext().fma_peak(nv.F64 if dtype == torch.float64 else nv.F32, int(fma_per_thread))
