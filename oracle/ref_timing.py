"""TEST INFRASTRUCTURE ONLY - times the UNMODIFIED reference Kalman filter class on host cores.

north_star: throughput is reported "next to the reference CPU filter timed on the same box's host cores in the same
run, with core count stated".  The class is loaded through oracle/ref_shim.py (live tree in the build container,
oracle/_ref on the GPU box) and stepped exactly as the hot path prescribes (kalman_filter.py:79-138,164-174):

    KF.set_measurements(imu, KF.get_odom(p, dp, contact, imu)); KF.predict(p, f); KF.update()

over the config-1 trajectory (synthetic stream of seed 0, 2,000 steps): best of `repeats` on one core, and one
trajectory per process on every core (SURVEY 8(d) "CPU baseline timing").  Used by bench.py's cpu_baseline leg only.
"""
from __future__ import annotations

import os
import time


def _one_trajectory(args):
    seed, n_steps = args
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    import numpy as np

    from optistate_b200.synth import make_stream
    from oracle import ref_shim

    KF_cls, _, ref_settings = ref_shim.load()
    s = make_stream(seed, n_steps)
    cols = {k: [s[k][t].reshape(-1, 1).copy() for t in range(n_steps)] for k in ("imu", "p", "dp", "contact", "f")}
    kf = KF_cls()
    kf.x = np.asarray(ref_settings.INITIAL_PARAMS.STARTING_STATE, float).reshape(12, 1).copy()
    kf.P = np.array(kf.Q, dtype=float).copy()
    t0 = time.perf_counter()
    for t in range(n_steps):
        imu = cols["imu"][t]
        kf.set_measurements(imu, kf.get_odom(cols["p"][t], cols["dp"][t], cols["contact"][t], imu))
        kf.predict(cols["p"][t], cols["f"][t])
        kf.update()
    dt = time.perf_counter() - t0
    return dt, float(kf.x[5, 0])


def available() -> bool:
    from oracle import ref_shim

    return ref_shim.available()


def time_reference_class(n_steps: int = 2000, repeats: int = 5, n_procs: int = 0):
    """Returns dict(steps_per_s_1core, steps_per_s_ncores, cores, x5_check) or None when the reference is not available."""
    if not available():
        return None
    best = min(_one_trajectory((0, n_steps))[0] for _ in range(repeats))
    out = {"reference_class_steps_per_s_1core": n_steps / best, "reference_class_trajectory": f"synthetic seed 0, {n_steps} steps (config 1)",
           "reference_class_x5_final": _one_trajectory((0, n_steps))[1]}
    cores = n_procs or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count() or 1)
    try:
        import multiprocessing as mp

        ctx = mp.get_context("spawn")  # the parent may hold a CUDA context: never fork it
        with ctx.Pool(cores) as pool:
            pool.map(_one_trajectory, [(100 + i, 50) for i in range(cores)])  # imports and first-call costs out of the way
            t0 = time.perf_counter()
            res = pool.map(_one_trajectory, [(i, n_steps) for i in range(cores)])
            wall = time.perf_counter() - t0
        out["reference_class_steps_per_s_ncores"] = cores * n_steps / wall
        out["reference_class_cores"] = cores
        out["reference_class_inloop_steps_per_s_ncores"] = sum(n_steps / r[0] for r in res)
    except Exception as e:  # noqa: BLE001 - a box that cannot spawn processes still reports the 1-core figure
        out["reference_class_ncores_error"] = str(e)[:200]
    return out


if __name__ == "__main__":
    import json

    print(json.dumps(time_reference_class()))
