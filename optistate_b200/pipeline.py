"""Double-buffered host -> device -> host pipeline around kf_batch for callers whose inputs and results live in host
memory: the upload of batch k+1 and the download of batch k-1 run on their own CUDA streams underneath the filter kernel
of batch k, so sustained end-to-end throughput is max(compute, copies) instead of their sum.

    pipe = KfHostPipeline(n_traj, n_steps, n_streams, dtype=torch.float64, labels=("truth", "nominal"))
    for batch in batches:                       # dicts of pinned host tensors: imu p dp contact f [truth nominal] Q R
        ticket = pipe.submit(batch)
        ...
        summary = pipe.result(ticket)           # pinned host tensor [52, n_traj], valid until the slot is reused
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from . import _native as nv
from .batch import kf_batch

_STREAM_CH = {"imu": 6, "p": 12, "dp": 12, "contact": 4, "f": 12, "truth": 12, "nominal": 12}


class KfHostPipeline:
    def __init__(self, n_traj: int, n_steps: int, n_streams: int, *, dtype: torch.dtype = torch.float64,
                 labels: Sequence[str] = ("truth", "nominal"), stream_offset: int = 0, n_slots: int = 2, device=None,
                 structure: str = "auto"):
        nv.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_traj, self.n_steps, self.n_streams, self.dtype = n_traj, n_steps, n_streams, dtype
        self.labels, self.stream_offset, self.n_slots = tuple(labels), int(stream_offset), n_slots
        self.structure = structure
        esz = 8 if dtype == torch.float64 else 4
        names = ["imu", "p", "dp", "contact", "f", *self.labels]
        self.slots = []
        for _ in range(n_slots):
            dev = {k: torch.empty((n_steps, _STREAM_CH[k], n_streams), dtype=dtype, device=self.device) for k in names}
            dev["Q"] = torch.empty((12, n_traj), dtype=dtype, device=self.device)
            dev["R"] = torch.empty((10, n_traj), dtype=dtype, device=self.device)
            out = {"summary": torch.empty((nv.SUMMARY_ROWS, n_traj), dtype=dtype, device=self.device),
                   "status": torch.zeros(n_traj, dtype=torch.int32, device=self.device),
                   "workspace": torch.empty(n_steps * 10 * n_streams * esz + 4 * n_streams + 4096, dtype=torch.uint8, device=self.device)}
            host = torch.empty((nv.SUMMARY_ROWS, n_traj), dtype=dtype).pin_memory()
            self.slots.append({"dev": dev, "out": out, "host": host, "ev_in": torch.cuda.Event(), "ev_k": torch.cuda.Event(),
                               "ev_out": torch.cuda.Event(), "busy": False})
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(device=self.device) for _ in range(3))
        self._next = 0

    def submit(self, host: Dict[str, torch.Tensor]) -> int:
        """Queues one batch (pinned host tensors) and returns its ticket; blocks only if all slots are still in flight."""
        i = self._next
        self._next = (self._next + 1) % self.n_slots
        slot = self.slots[i]
        if slot["busy"]:
            slot["ev_out"].synchronize()  # the previous result of this slot must have left the device
        dev = slot["dev"]
        with torch.cuda.stream(self.s_in):
            if slot["busy"]:
                self.s_in.wait_event(slot["ev_k"])  # do not overwrite inputs a kernel may still be reading
            for k, d in dev.items():
                d.copy_(host[k], non_blocking=True)
            slot["ev_in"].record(self.s_in)
        with torch.cuda.stream(self.s_k):
            self.s_k.wait_event(slot["ev_in"])
            kf_batch(dev["imu"], dev["p"], dev["dp"], dev["contact"], dev["f"], Q=dev["Q"], R=dev["R"], n_traj=self.n_traj,
                     dtype=self.dtype, stream_offset=self.stream_offset, truth=dev.get("truth"), nominal=dev.get("nominal"),
                     outputs=("summary",), out=slot["out"], q_kind=nv.MAT_DIAG_PER, r_kind=nv.MAT_DIAG_PER, device=self.device,
                     structure=self.structure)
            slot["ev_k"].record(self.s_k)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["ev_k"])
            slot["host"].copy_(slot["out"]["summary"], non_blocking=True)
            slot["ev_out"].record(self.s_out)
        slot["busy"] = True
        return i

    def result(self, ticket: int) -> torch.Tensor:
        """Pinned host summary [52, n_traj] of a submitted batch (waits for its download)."""
        slot = self.slots[ticket]
        slot["ev_out"].synchronize()
        return slot["host"]

    def drain(self):
        for slot in self.slots:
            if slot["busy"]:
                slot["ev_out"].synchronize()
