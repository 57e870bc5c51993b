import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from optistate_b200 import kf_batch
from optistate_b200.synth import make_streams
S, T = 128, 40
st = make_streams(range(S), T)
outs = ("x_steps", "x_model_steps", "p_world_steps", "z_steps", "p_trace", "nis", "P_ckpt")
os.environ["OPTISTATE_KF_PACKED"] = "1"
a = kf_batch(st["imu"], st["p"], st["dp"], st["contact"], st["f"], dtype=torch.float32, outputs=outs, ckpt_every=1)
os.environ["OPTISTATE_KF_PACKED"] = "0"
b = kf_batch(st["imu"], st["p"], st["dp"], st["contact"], st["f"], dtype=torch.float32, outputs=outs, ckpt_every=1)
for name in ("z_steps", "p_world_steps", "x_model_steps", "P_ckpt", "x_steps", "p_trace_steps", "nis_steps"):
    A, B = a.tensors[name].cpu().numpy(), b.tensors[name].cpu().numpy()
    d = np.argwhere(A != B)
    print(name, "n_diff", len(d), "first", d[0] if len(d) else None, (A[tuple(d[0])], B[tuple(d[0])]) if len(d) else "")
