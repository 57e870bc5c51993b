import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from optistate_b200 import kf_batch
from optistate_b200.synth import make_streams
S, T = 1024, 2000
st = make_streams(range(S), T)
dev = {k: torch.from_numpy(v).cuda() for k, v in st.items()}
for algo in ("sequential", "joint"):
    for outs in (("x_steps", "p_trace", "k_gain"), ("x_steps",), ("x_final",)):
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = kf_batch(dev["imu"], dev["p"], dev["dp"], dev["contact"], dev["f"], outputs=outs, algo=algo)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(algo, outs, f"{best:.2f} ms  {S*T/best/1e-3:.3e} steps/s  {best*1e-3/T*1.965e9:.0f} cycles/step", flush=True)
