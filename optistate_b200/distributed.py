"""Multi-GPU plumbing: trajectories are independent, so they are sharded in contiguous blocks, one block per
rank (one process per GPU), with NO traffic between GPUs inside the time loop.  The only collective is the
all-gather of the per-trajectory summaries (and optionally final states) after the loop - NCCL over NVLink /
NVSwitch on the GPU box, gloo in the CPU tests of this host logic.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) block of trajectories owned by `rank`; the first n_total % world_size ranks get one more."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(n_total: int, world_size: int):
    return [shard_range(n_total, world_size, r)[1] - shard_range(n_total, world_size, r)[0] for r in range(world_size)]


def stream_exchange_group(ranks=None):
    """A process group for KfHostPipeline's exchange of base-stream slices (`shared_streams_group`): NCCL with HIGH-PRIORITY
    streams, so that the all-gather of batch k+1's slices is scheduled as soon as a block of batch k's filter kernel retires
    instead of queueing behind the whole kernel (the filter grid keeps every SM occupied for the length of a batch; on a
    normal-priority stream the exchange ran only after it - measured: profiles/r2_e2e_upload_modes.txt).  gloo groups (CPU
    tests of the host logic) have no such option and are returned as they are."""
    if dist.get_backend() != "nccl":
        return dist.new_group(ranks=ranks)
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    return dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)


def gather_columns(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gathers per-trajectory columns: local [C, n_local] on every rank -> [C, n_total] on every rank, in
    trajectory order.  Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[-1] != sizes[dist.get_rank(group)]:
        raise ValueError(f"local shard has {local.shape[-1]} trajectories, expected {sizes[dist.get_rank(group)]}")
    pad = max(sizes)
    C = local.shape[0]
    if pad == local.shape[-1]:
        send = local.contiguous()
    else:
        send = local.new_zeros((C, pad))
        send[:, : local.shape[-1]] = local
    recv = local.new_empty((world * C, pad))  # one [C, pad] block per rank
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, C, pad)
    out = local.new_empty((C, n_total))
    col = 0
    for r in range(world):
        out[:, col : col + sizes[r]] = recv[r, :, : sizes[r]]
        col += sizes[r]
    return out


def kf_batch_sharded(streams: Dict[str, torch.Tensor], n_total: int, *, Q=None, R=None, x0=None, member_offset: int = 0,
                     gather=True, group=None, **kw):
    """Filters this rank's contiguous block of `n_total` trajectories and (optionally) all-gathers the summaries.

    streams: dict with imu, p, dp, contact, f (and optional truth, nominal), replicated on every rank ([T, C, S]).
    Q, R, x0: either shared, or per-trajectory arrays for the LOCAL block ([C, n_local]).
    Trajectory i (global id) reads stream (i + member_offset) % S, so the shard boundary does not change the mapping.
    gather: True / "nccl" = NCCL all-gather after the kernel; a peer.PeerSummary = the all-gather fused into the filter
    kernel (NVLink peer stores, see peer.py); False = no gather.
    Returns (KfBatchResult of the local block, gathered summary [52, n_total] or None).
    """
    from .batch import kf_batch

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    begin, end = shard_range(n_total, world, rank)
    outputs = tuple(kw.pop("outputs", ("summary",)))
    if gather is not False and "summary" not in outputs:
        outputs = outputs + ("summary",)
    fused = gather if hasattr(gather, "cfg") else None
    if fused is not None:
        if fused.n_total != n_total:
            raise ValueError("PeerSummary was set up for another trajectory count")
        fused.wait()  # every rank is done with the previous contents of the shared array
        kw["summary_peers"] = fused
    res = kf_batch(streams["imu"], streams["p"], streams["dp"], streams["contact"], streams["f"], x0=x0, Q=Q, R=R,
                   n_traj=end - begin, stream_offset=begin + member_offset, truth=streams.get("truth"),
                   nominal=streams.get("nominal"), outputs=outputs, **kw)
    gathered: Optional[torch.Tensor] = None
    if fused is not None:
        fused.wait()  # all kernels have finished: every copy is complete
        gathered = fused.tensor
    elif gather:
        gathered = gather_columns(res.summary, n_total, group=group)
    return res, gathered
