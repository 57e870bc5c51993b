"""Peer memory for the fused all-gather of the per-trajectory summaries (one process per GPU, one NVSwitch box).

Every rank owns one copy of the job-wide summary array [52, n_total] in a dedicated device allocation and maps the
copies of all other ranks into its own address space (cudaIpc through the C ABI: optistate_kf_peer_*).  The filter
kernel then stores each summary value of its trajectories to the same element of all copies - plain NVLink peer
stores issued as trajectories finish, while the other warps of the GPU are still filtering - so the collective
costs no extra pass over the data, no staging buffers and no extra kernel.  NCCL is left with what it is needed
for: one tiny all-reduce as the "all kernels have finished" barrier.

The NCCL all-gather (distributed.gather_columns) stays available; it is what the fused path is measured against.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _native as nv
from .distributed import shard_range


class PeerSummary:
    """Job-wide summary array replicated on every GPU; `tensor` is this rank's copy ([52, n_total]),
    `local` the columns this rank's kernel fills."""

    def __init__(self, n_total: int, dtype: torch.dtype = torch.float64, group=None, device=None):
        nv.require_cuda()
        if not dist.is_initialized():
            raise RuntimeError("PeerSummary needs an initialised process group (one process per GPU)")
        ext = nv.ext()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world - 1 > ext.MAX_PEERS:
            raise ValueError(f"at most {ext.MAX_PEERS + 1} GPUs (one box)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype, self.n_total = dtype, int(n_total)
        self.begin, self.end = shard_range(self.n_total, self.world, self.rank)
        self.n_local = self.end - self.begin
        esz = torch.empty((), dtype=dtype).element_size()
        # set-up failures (no peer access between two GPUs, IPC refused by the container, ...) are agreed on by all
        # ranks before anyone raises, so that the collectives of the caller stay matched
        self._raw, self._peers, err, handle = None, [], None, None
        try:
            self._raw = ext.peer_alloc(nv.SUMMARY_ROWS * self.n_total * esz, self.device.index)
            handle = bytes(ext.peer_export(self._raw))
        except Exception as e:  # noqa: BLE001 - reported to every rank below
            err = f"rank {self.rank}: {e}"
        infos = [None] * self.world
        dist.all_gather_object(infos, (handle, err), group=group)
        self._raise_if_any([e for _, e in infos])
        try:
            self._peers = [ext.peer_open(h, self.device.index) for r, (h, _) in enumerate(infos) if r != self.rank]
        except Exception as e:  # noqa: BLE001
            err = f"rank {self.rank}: {e}"
        errs = [None] * self.world
        dist.all_gather_object(errs, err, group=group)
        self._raise_if_any(errs)
        self.tensor = self._raw.view(dtype).view(nv.SUMMARY_ROWS, self.n_total)
        self.local = self.tensor[:, self.begin:self.end]
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.wait()  # nobody launches before every mapping exists

    def _raise_if_any(self, errs) -> None:
        errs = [e for e in errs if e]
        if not errs:
            return
        ext = nv.ext()
        for a in self._peers:
            ext.peer_close(int(a), self.device.index)
        self._peers, self._raw = [], None
        raise RuntimeError("peer memory unavailable: " + "; ".join(errs))

    def cfg(self) -> dict:
        """Integer settings the binding turns into OptiKfDesc.summary_ld / summary_peers."""
        c = {"summary_ld": self.n_total, "summary_col0": self.begin, "n_summary_peers": len(self._peers)}
        c.update({f"summary_peer{k}": int(a) for k, a in enumerate(self._peers)})
        return c

    def wait(self) -> None:
        """Stream-ordered barrier across the ranks: once it has run on this rank's stream, the kernels every rank
        queued before its own wait() have finished, i.e. this rank's copy of the array is complete (and, called
        before a launch, every rank is done reading the previous contents)."""
        dist.all_reduce(self._flag, group=self.group)

    def close(self) -> None:
        if self._raw is None:
            return
        ext = nv.ext()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for a in self._peers:
            ext.peer_close(int(a), self.device.index)
        self._peers = []
        dist.barrier(group=self.group)  # every mapping of this rank's copy is gone before it is released
        self.tensor = self.local = self._raw = None
