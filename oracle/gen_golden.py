"""TEST INFRASTRUCTURE ONLY - mints tests/golden/*.npz by running the UNMODIFIED reference class.

Run in the build container (where /root/reference exists):   python -m oracle.gen_golden
For every case of oracle/cases.py the reference `Kalman_Filter` (loaded through oracle/ref_shim.py)
is stepped exactly as SURVEY.md 3.2 describes and its observable outputs are stored, sub-sampled in
time for the long cases.  q_r_pkl.npz holds the 22 diagonal values of the reference's only data
fixture (data_collection/trajectories/Q_R.pkl) so that the stress case can be rebuilt on the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import cases, ref_shim  # noqa: E402


def input_digest(stream) -> str:
    h = hashlib.sha256()
    for k in ("imu", "p", "dp", "contact", "f"):
        h.update(np.ascontiguousarray(stream[k], dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
    with open(os.path.join(ref_shim.REF_ROOT, "data_collection", "trajectories", "Q_R.pkl"), "rb") as fh:
        Q, R = pickle.load(fh)
    assert np.count_nonzero(Q - np.diag(np.diag(Q))) == 0 and np.count_nonzero(R - np.diag(np.diag(R))) == 0
    np.savez(os.path.join(cases.GOLDEN_DIR, "q_r_pkl.npz"), q_diag=np.diag(Q), r_diag=np.diag(R))

    for name in cases.ALL_CASES:
        stream, kw, stride = cases.build(name)
        T = stream["imu"].shape[0]
        every = 1000 if T >= 2000 else max(T // 4, 1)
        ref = ref_shim.run_reference(stream, x0=kw["x0"], P0=kw["P0"], Q=kw["Q"], R=kw["R"],
                                     p_checkpoint_every=every, mode=kw["model"])
        steps = np.arange(stride - 1, T, stride)
        ck = sorted(ref["P_ckpt"])
        np.savez(
            os.path.join(cases.GOLDEN_DIR, name + ".npz"),
            steps=steps, x=ref["x"][steps], x_model=ref["x_model"][steps], z=ref["z"][steps],
            p_world=ref["p_world"][steps], p_trace=ref["p_trace"][steps], k_gain=ref["k_gain"][steps],
            ckpt_steps=np.array(ck), P_ckpt=np.stack([ref["P_ckpt"][k] for k in ck]),
            P_final=ref["P_final"], K_last=ref["K_last"], x_absmax=np.abs(ref["x"]).max(axis=0),
            input_sha256=np.array(input_digest(stream)),
        )
        print(f"{name}: T={T} stored {len(steps)} steps, {len(ck)} P checkpoints")


if __name__ == "__main__":
    main()
