"""GRU input preparation on the device: the step right after the filter.

    rows    = assemble_features(res, streams)                # [N, T, 60]  driver rows, Kalman_to_Training.py:245-254
    lo, hi  = min_max(rows.reshape(-1, 60))                  # gru_train.py:56-63
    windows = normalized_windows(rows.reshape(-1, 60), lo, hi, latent, n_groups=N, seq_len=10)   # gru_train.py:108-111,180-192

Everything stays in HBM: filter outputs -> feature rows -> float32 windows the reference's `RNN.forward` consumes.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as nv


def _code(dtype):
    return nv.F64 if dtype == torch.float64 else nv.F32


def assemble_features(x_steps: torch.Tensor, p_world_steps: torch.Tensor, imu, f, dp, imu_acc=None, *, stream_index=None,
                      stream_offset: int = 0) -> torch.Tensor:
    """x_steps, p_world_steps: [T, 12, N] (kf_batch outputs); imu [T,6,S], f [T,12,S], dp [T,12,S], imu_acc [T,6,S] or None.
    Returns rows [N, T, 60] = [x | imu_acc | f | p_world | dp | imu] per (trajectory, step), in the filter's dtype."""
    nv.require_cuda()
    dtype, dev = x_steps.dtype, x_steps.device
    T, _, N = x_steps.shape
    S = imu.shape[2]
    as_dev = lambda a: None if a is None else torch.as_tensor(a).to(device=dev, dtype=dtype).contiguous()  # noqa: E731
    tensors = {"x_steps": x_steps.contiguous(), "p_world_steps": p_world_steps.contiguous(), "imu": as_dev(imu), "f": as_dev(f), "dp": as_dev(dp),
               "rows": torch.empty((N, T, 60), dtype=dtype, device=dev)}
    if imu_acc is not None:
        tensors["imu_acc"] = as_dev(imu_acc)
    if stream_index is not None:
        from .batch import _checked_stream_index

        tensors["stream_index"] = _checked_stream_index(stream_index, N, S, dev)
    with torch.cuda.device(dev):
        nv.check(nv.ext().kf_features(_code(dtype), N, T, S, int(stream_offset), tensors), "optistate_kf_features")
    return tensors["rows"]


def min_max(rows: torch.Tensor):
    """Per-column (min, max) of a [R, C] matrix on the device."""
    nv.require_cuda()
    rows = rows.contiguous()
    lo = torch.empty(rows.shape[1], dtype=rows.dtype, device=rows.device)
    hi = torch.empty_like(lo)
    nv.check(nv.ext().kf_minmax(rows, lo, hi), "optistate_kf_minmax")
    return lo, hi


def normalized_windows(rows: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, latent: Optional[torch.Tensor] = None, *, n_groups: int = 1,
                       seq_len: int = 10) -> torch.Tensor:
    """(rows - lo) / (hi - lo), latent appended, sliding windows inside each of the n_groups equal row groups, float32:
    [n_groups, R/n_groups - seq_len + 1, seq_len, C + n_latent].  n_groups = 1 reproduces the reference's windows over the
    concatenated datasets; n_groups = N keeps windows inside one trajectory."""
    nv.require_cuda()
    rows = rows.contiguous()
    R, C = rows.shape
    rpg = R // n_groups
    n_lat = 0 if latent is None else latent.shape[1]
    out = torch.empty((n_groups, rpg - seq_len + 1, seq_len, C + n_lat), dtype=torch.float32, device=rows.device)
    nv.check(nv.ext().kf_windows(rows, latent, lo, hi, int(n_groups), int(seq_len), out), "optistate_kf_windows")
    return out
