"""Developer helper (GPU box): QPs/s of the batched force MPC.  usage: python tools/mpc_rate.py [n_problems] [trot | stand | three]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optistate_b200.mpc import mpc_forces  # noqa: E402
from tests import mpc_cases  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
x, ref, p, c = (torch.from_numpy(a).cuda() for a in mpc_cases.batch(4096, seed=1))
rep = n // 4096
x, ref, p, c = x.repeat(1, rep), ref.repeat(1, 1, rep), p.repeat(1, rep), c.repeat(1, rep)
if len(sys.argv) > 2 and sys.argv[2] == "trot":  # the gait of the recordings: diagonal leg pairs alternate
    c = torch.where((torch.arange(c.shape[1], device=c.device) % 2 == 0)[None, :], torch.tensor([1.0, 0, 0, 1], device=c.device, dtype=c.dtype)[:, None],
                    torch.tensor([0, 1.0, 1, 0], device=c.device, dtype=c.dtype)[:, None]).contiguous()
if len(sys.argv) > 2 and sys.argv[2] in ("stand", "three"):  # all four legs in stance / three (a walk)
    pat = [1.0, 1, 1, 1] if sys.argv[2] == "stand" else [1.0, 1, 0, 1]
    c = torch.tensor(pat, device=c.device, dtype=c.dtype)[:, None].repeat(1, c.shape[1]).contiguous()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    forces, status = mpc_forces(x, ref, p, c)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
it = (status >> 8).double()
print(f"{x.shape[1]} QPs in {best:.2f} ms = {x.shape[1] / best / 1e-3:.3e} QP/s; interior-point iterations mean {it.mean():.1f} max {int(it.max())}; "
      f"unpolished {int((status & 2).ne(0).sum())}, ipm-limit {int((status & 1).ne(0).sum())}")
