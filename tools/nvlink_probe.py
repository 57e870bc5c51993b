"""Developer helper (GPU box with >= 2 GPUs): ONE process, the filter kernel on cuda:0 storing its summaries into a second
copy of the job-wide array that lives on cuda:1 (the same peer stores PeerSummary sets up across processes), so that a
single-process ncu capture can read the NVLink byte counters of the fused all-gather.
usage: ncu --metrics regex:nvl.*bytes -k regex:kf_seq_tma python tools/nvlink_probe.py [N] [T]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from optistate_b200 import kf_batch  # noqa: E402
from optistate_b200 import _native as nv  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
S = 1024
assert torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1)
torch.cuda.set_device(0)
st = make_streams(range(S), T)
d = {k: torch.from_numpy(st[k]).cuda() for k in ("imu", "p", "dp", "contact", "f", "truth")}
q, r = bench.mc_noise(0, N, S)


class TwoGpuPeers:
    """Attributes of optistate_b200.peer.PeerSummary; rank 0 of 2, the peer copy on cuda:1."""

    def __init__(self):
        self.device, self.dtype, self.n_total, self.begin, self.end, self.n_local = torch.device("cuda", 0), torch.float64, 2 * N, 0, N, N
        self.tensor = torch.zeros((nv.SUMMARY_ROWS, 2 * N), dtype=torch.float64, device="cuda:0")
        self.local = self.tensor[:, :N]
        self.remote = torch.zeros((nv.SUMMARY_ROWS, 2 * N), dtype=torch.float64, device="cuda:1")
        import ctypes

        rt = ctypes.CDLL("libcudart.so.12")
        torch.cuda.set_device(0)
        rc = rt.cudaDeviceEnablePeerAccess(1, 0)  # cuda:0 may store into allocations of cuda:1 (what cudaIpcOpenMemHandle sets up across processes)
        assert rc in (0, 704), rc                 # 704 = already enabled

    def cfg(self):
        return {"summary_ld": self.n_total, "summary_col0": 0, "n_summary_peers": 1, "summary_peer0": self.remote.data_ptr()}


peers = TwoGpuPeers()
for _ in range(2):
    res = kf_batch(d["imu"], d["p"], d["dp"], d["contact"], d["f"], Q=q, R=r, n_traj=N, truth=d["truth"], outputs=("summary",),
                   q_kind=nv.MAT_DIAG_PER, r_kind=nv.MAT_DIAG_PER, summary_peers=peers)
torch.cuda.synchronize()
same = torch.equal(peers.remote[:, :N].to("cuda:0"), peers.tensor[:, :N])
print(f"nvlink_probe: N={N} T={T} peer copy identical: {same}; algorithmic peer-store bytes per launch: {nv.SUMMARY_ROWS * 8 * N}", flush=True)
assert same and float(peers.tensor[:, :N].abs().max()) > 0
