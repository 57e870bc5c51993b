"""GPU tests of the fused all-gather of the summaries: the filter kernels store every summary value into this rank's
columns of the job-wide [52, n_total] array and into every peer copy.

On one GPU the peer copy is a second array on the same device (a loop-back stand-in for PeerSummary - the kernel
cannot tell); with >= 2 GPUs the real thing runs as two NCCL ranks and is compared with the NCCL all-gather."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from optistate_b200 import kf_batch
from optistate_b200 import _native as nv
from optistate_b200.settings import INITIAL_PARAMS
from optistate_b200.synth import make_streams, monte_carlo_noise

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SENTINEL = -777.0


class LoopbackPeers:
    """Same attributes as optistate_b200.peer.PeerSummary; the 'peers' are further arrays on the same GPU."""

    def __init__(self, n_total, begin, n_local, dtype, n_peers=2):
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype, self.n_total, self.begin, self.end, self.n_local = dtype, n_total, begin, begin + n_local, n_local
        self.tensor = torch.full((nv.SUMMARY_ROWS, n_total), SENTINEL, dtype=dtype, device=self.device)
        self.local = self.tensor[:, self.begin:self.end]
        self.copies = [torch.full((nv.SUMMARY_ROWS, n_total), SENTINEL, dtype=dtype, device=self.device) for _ in range(n_peers)]

    def cfg(self):
        c = {"summary_ld": self.n_total, "summary_col0": self.begin, "n_summary_peers": len(self.copies)}
        c.update({f"summary_peer{k}": t.data_ptr() for k, t in enumerate(self.copies)})
        return c


def _inputs(S, T, N, dtype):
    st = make_streams(range(S), T)
    d = {k: torch.from_numpy(st[k]).to(dtype).cuda() for k in ("imu", "p", "dp", "contact", "f", "truth")}
    q, r = monte_carlo_noise(np.arange(N), np.diag(INITIAL_PARAMS.Q).copy(), np.diag(INITIAL_PARAMS.R).copy())
    return d, torch.from_numpy(q).to(dtype).cuda(), torch.from_numpy(r).to(dtype).cuda()


@pytest.mark.parametrize("dtype,path", [(torch.float64, "tma"), (torch.float32, "tma"), (torch.float64, "direct"),
                                        (torch.float32, "direct"), (torch.float64, "joint")])
def test_summary_is_stored_to_every_copy_at_the_right_columns(dtype, path):
    S, T, N = 64, 40, 256
    d, q, r = _inputs(S, T, N, dtype)
    kw = dict(Q=q, R=r, n_traj=N, dtype=dtype, truth=d["truth"], outputs=("summary",), q_kind=nv.MAT_DIAG_PER, r_kind=nv.MAT_DIAG_PER)
    if path == "direct":
        kw["stream_index"] = (torch.arange(N, dtype=torch.int32) * 7 + 64) % S  # a gather: the direct-load kernel
    else:
        kw["stream_offset"] = 64
    if path == "joint":
        kw["algo"] = "joint"
    plain = kf_batch(d["imu"], d["p"], d["dp"], d["contact"], d["f"], **kw)
    n_total, begin = 3 * N + 64, N + 64
    peers = LoopbackPeers(n_total, begin, N, dtype)
    fused = kf_batch(d["imu"], d["p"], d["dp"], d["contact"], d["f"], summary_peers=peers, **kw)
    torch.cuda.synchronize()
    assert fused.algo == plain.algo
    assert torch.equal(fused.summary, plain.summary)  # bit-identical values, only the destination differs
    assert fused.summary.data_ptr() == peers.tensor[:, begin:].data_ptr()
    for arr in [peers.tensor] + peers.copies:
        assert torch.equal(arr[:, begin:begin + N], plain.summary)
        assert (arr[:, :begin] == SENTINEL).all() and (arr[:, begin + N:] == SENTINEL).all()  # nobody else's columns touched


def test_odd_row_stride_takes_the_one_trajectory_fp32_kernel():
    """The packed FP32 kernel stores pairs of summary values as one float2, which needs an even row stride."""
    S, T, N = 64, 20, 128
    d, q, r = _inputs(S, T, N, torch.float32)
    kw = dict(Q=q, R=r, n_traj=N, dtype=torch.float32, truth=d["truth"], outputs=("summary",), q_kind=nv.MAT_DIAG_PER, r_kind=nv.MAT_DIAG_PER)
    plain = kf_batch(d["imu"], d["p"], d["dp"], d["contact"], d["f"], packed=False, **kw)
    peers = LoopbackPeers(2 * N + 1, N, N, torch.float32, n_peers=1)
    fused = kf_batch(d["imu"], d["p"], d["dp"], d["contact"], d["f"], summary_peers=peers, **kw)
    torch.cuda.synchronize()
    assert torch.equal(fused.summary, plain.summary) and torch.equal(peers.copies[0][:, N:2 * N], plain.summary)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank = int(sys.argv[3]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=2, device_id=torch.device("cuda", rank))
from optistate_b200.distributed import kf_batch_sharded, shard_range
from optistate_b200.peer import PeerSummary
from optistate_b200.synth import make_streams, monte_carlo_noise
from optistate_b200 import _native as nv
S, T, n_total = 64, 50, 64 * 9 + 2   # uneven shards
st = make_streams(range(S), T)
streams = {k: torch.from_numpy(st[k]).cuda() for k in ("imu", "p", "dp", "contact", "f", "truth")}
b, e = shard_range(n_total, 2, rank)
import numpy as np
from optistate_b200.settings import INITIAL_PARAMS
q, r = monte_carlo_noise(np.arange(n_total), np.diag(INITIAL_PARAMS.Q).copy(), np.diag(INITIAL_PARAMS.R).copy())
kw = dict(Q=torch.from_numpy(q[:, b:e].copy()).cuda(), R=torch.from_numpy(r[:, b:e].copy()).cuda(), q_kind=nv.MAT_DIAG_PER, r_kind=nv.MAT_DIAG_PER)
_, via_nccl = kf_batch_sharded(streams, n_total, gather=True, **kw)
peers = PeerSummary(n_total, torch.float64)
for _ in range(3):  # reuse of the shared array across calls
    res, fused = kf_batch_sharded(streams, n_total, gather=peers, **kw)
torch.cuda.synchronize()
assert fused.shape == (52, n_total) and torch.equal(fused, via_nccl), (fused - via_nccl).abs().max()
assert torch.equal(res.summary, via_nccl[:, b:e])
peers.close(); dist.barrier(); dist.destroy_process_group(); print("ok")
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
def test_fused_all_gather_equals_nccl_all_gather_on_two_gpus(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0 and "ok" in out, err[-3000:]
