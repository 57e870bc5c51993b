"""Developer helper (GPU box): which problems of the all-patterns test batch does the dual active-set kernel hand to the interior point?
Times the batch per contact pattern (a sub-batch that contains such a problem shows the ~3 ms floor of one interior-point solve) and
lists, for the slow patterns, the problems whose status carries interior-point iteration counts."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optistate_b200.mpc import mpc_forces  # noqa: E402
from tests import mpc_cases  # noqa: E402

x, ref, p, c = (torch.from_numpy(a).cuda() for a in mpc_cases.batch(4096, seed=1))
key = (c != 0).to(torch.int64)
code = key[0] * 8 + key[1] * 4 + key[2] * 2 + key[3]
for pat in sorted(set(code.tolist())):
    idx = torch.nonzero(code == pat)[:, 0]
    if idx.numel() == 0:
        continue
    xs, rs, ps, cs = x[:, idx].contiguous(), ref[:, :, idx].contiguous(), p[:, idx].contiguous(), c[:, idx].contiguous()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f, st = mpc_forces(xs, rs, ps, cs, max_free_legs=4)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    it = (st >> 8)
    cold = mpc_forces(xs, rs, ps, cs, max_free_legs=4, max_changes=100000)[1]
    print(f"pattern {pat:04b}: {idx.numel():5d} problems, {best:6.2f} ms; changes mean {it.double().mean():5.1f} max {int(it.max()):3d}; "
          f"status bits {sorted(set((st & 0xff).tolist()))}; contact values {sorted(set(cs.flatten().tolist()))}", flush=True)
