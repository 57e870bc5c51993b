#!/usr/bin/env python
"""Headline benchmark: Kalman-filter trajectory-steps/sec on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--dtype f64|f32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): the Monte-Carlo noise sweep of BASELINE.json configs[2]/[3] - 1,048,576 trajectories
x 1,000 steps PER GPU over 1,024 shared synthetic base streams, per-member diagonal Q/R perturbations, outputs =
per-trajectory summaries (final x, diag P, RMSE vs the label stream, RMS deviation from the nominal member, mean
NIS, ...) - run in FP64 by default because the north-star target is stated on the FP64 FMA roofline with 1e-9
parity (--dtype f32 gives configs[2] verbatim).  One "step" of the bench = one full pass of the hot path over
that batch: measurement pre-pass + the filter kernel; when N > 1 the all-gather of the summaries is fused into the
filter kernel (NVLink peer stores, optistate_b200/peer.py; --gather nccl runs the NCCL all-gather instead).

  value      whole-job trajectory-steps/s, inputs resident in HBM, timed with CUDA events, max over ranks
  e2e        same through the public kf_batch() call with pinned HOST buffers: H2D of streams + per-member noise
             and D2H of the summaries inside the timed region
  roofline   FMA roofline: achieved = 6,800 algorithmic flops/step x steps/s; peak = FP64 (FP32)
             FMA issue peak MEASURED in this run by optistate_fma_peak; executed-flop figures alongside
  cpu_baseline  the oracle's C port of the reference filter on this box's host cores, bounded sample

--impl reference times the CPU arm (oracle C port, all host threads) on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_ALGORITHMIC = 6800.0  # per trajectory-step, reference operand order, structural zeros skipped
# flops the streamed SEQUENTIAL kernel actually executes per trajectory-step in this workload (summary on, two label
# streams), counted by ncu in the SASS of its time loop (profiles/r1_ncu_bench_kernel_*_hotloop.txt):
#   FP64: 1135 DFMA + 321 DMUL + 99 DADD;  FP32 (two trajectories per thread): (1051 FFMA2 + 302 FMUL2 + 85 FADD2) x 2 lanes
#   + 54 FFMA + 30 FMUL + 8 FADD per PAIR of trajectories
FLOPS_EXECUTED = {"f64": 2 * 1135 + 321 + 99, "f32": (2 * 2 * 1051 + 2 * 302 + 2 * 85 + 2 * 54 + 30 + 8) / 2}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel on the default workload, from the ncu
# captures summarised in profiles/r1_ncu_bench_kernel_*_metrics.txt; None for any other workload shape
TRAFFIC_BYTES = {("f64", 1 << 20, 1000, 1024): 2.724e9 + 0.433e9, ("f32", 1 << 20, 1000, 1024): 1.230e9 + 0.212e9}
# algorithmic bytes of one launch: every base-stream channel read once (p, f, z, two label streams = 58 scalars per
# stream-step) + 22 noise scalars in and 52 summary scalars out per trajectory
def algorithmic_bytes(a, esz):
    return (58 * a.T * a.streams + (22 + 52) * a.traj_per_gpu) * esz
METRIC = "kf_trajectory_steps_per_sec"
UNIT = "trajectory-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--traj-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N > 1: all-gather of the summaries fused into the filter kernel (NVLink peer stores) or NCCL after it")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": f"monte-carlo noise sweep: {a.traj_per_gpu} trajectories x {a.T} steps per GPU, {a.streams} shared base "
                    f"streams, per-member diagonal Q/R, summary outputs (BASELINE configs[2]/[3] shape, {a.dtype})",
        "trajectories_per_gpu": a.traj_per_gpu, "trajectories_total": a.traj_per_gpu * world, "steps_per_trajectory": a.T,
        "base_streams": a.streams, "sharding": f"contiguous blocks x{world}, all-gather of summaries" if world > 1 else "single GPU",
        "l2": "no explicit flush: per-iteration inputs (streams + per-member noise) exceed the 126 MB L2",
    }


Q_DIAG = np.array([0.01, 0.01, 0.01, 0.01, 0.0001, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.0001])  # settings.py:28
R_DIAG = np.full(10, 0.01)  # settings.py:30
BLOCK = 1 << 16


def mc_noise(first_member: int, count: int, n_streams: int):
    """Vectorised Monte-Carlo noise (same law as synth.monte_carlo_noise: 10**U(-0.5,0.5) per diagonal entry; the first
    pass over the streams stays nominal).  Drawn in blocks of 65,536 members so it is a function of the member id."""
    q = np.empty((12, count))
    r = np.empty((10, count))
    done = 0
    while done < count:
        m = first_member + done
        blk, off = divmod(m, BLOCK)
        n = min(BLOCK - off, count - done)
        rng = np.random.default_rng([10**6, blk])
        u = rng.uniform(-0.5, 0.5, (BLOCK, 22))[off:off + n]
        q[:, done:done + n] = (Q_DIAG[None, :] * 10.0 ** u[:, :12]).T
        r[:, done:done + n] = (R_DIAG[None, :] * 10.0 ** u[:, 12:]).T
        done += n
    ids = first_member + np.arange(count)
    nominal = ids < n_streams
    q[:, nominal] = Q_DIAG[:, None]
    r[:, nominal] = R_DIAG[:, None]
    return q, r


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(a, seconds_target=12.0, n_threads=0):
    """The oracle's C port of the reference filter (oracle/kf_oracle.c) on a bounded sample of the same workload:
    n_sample trajectories x T steps over the same base streams with Monte-Carlo noise.  Returns (steps/s, dict)."""
    from oracle import c_oracle
    from optistate_b200.synth import make_streams

    threads = n_threads or c_oracle.max_threads()
    S = min(a.streams, 64)
    st = make_streams(range(S), a.T)
    q, r = mc_noise(0, threads, S)
    t0 = time.perf_counter()
    c_oracle.run(st, threads, Q=q, R=r, n_threads=threads, want=("x_final",))  # calibration: one trajectory per thread
    dt = time.perf_counter() - t0
    per_thread = max(1, int(seconds_target / max(dt, 1e-3)))
    n = threads * min(per_thread, 64)
    q, r = mc_noise(0, n, S)
    t0 = time.perf_counter()
    c_oracle.run(st, n, Q=q, R=r, n_threads=threads, want=("x_final",))
    dt = time.perf_counter() - t0
    return n * a.T / dt, {"cores": threads, "kind": "port", "sample": f"{n} trajectories x {a.T} steps ({dt:.1f} s), C port of the reference filter "
                          f"(oracle/kf_oracle.c, pthreads, one block of trajectories per thread)"}


def numpy_port_rate(steps=400):
    """Interpreter-bound NumPy restatement (what the reference's own loop costs per step, one core)."""
    from oracle import kf_numpy
    from optistate_b200.synth import make_stream

    s = make_stream(0, steps)
    t0 = time.perf_counter()
    kf_numpy.run(s)
    return steps / (time.perf_counter() - t0)


def run_reference(a):
    """CPU arm: oracle C port with all host threads; a step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    from optistate_b200.synth import make_streams

    threads = c_oracle.max_threads()
    S = min(a.streams, 64)
    st = make_streams(range(S), a.T)
    n = threads * 16
    q, r = mc_noise(0, n, S)
    fn = lambda: c_oracle.run(st, n, Q=q, R=r, n_threads=threads, want=("x_final",))  # noqa: E731
    t0 = time.perf_counter()
    fn()
    first = time.perf_counter() - t0
    if first * (a.steps + a.warmup) > 240:  # keep the whole arm within a few minutes
        n = max(threads, int(n * 240 / (first * (a.steps + a.warmup))))
        q, r = mc_noise(0, n, S)
    for _ in range(max(a.warmup - 1, 0)):
        fn()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    dt = time.perf_counter() - t0
    value = n * a.T * a.steps / dt
    sample = f"{n} trajectories x {a.T} steps per step over {S} base streams, Monte-Carlo Q/R"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "numpy_port_steps_per_s_1core": numpy_port_rate()}
    print(json.dumps(line), flush=True)


def run_native(a):
    import torch
    import torch.distributed as dist

    from optistate_b200 import fma_peak, kf_batch
    from optistate_b200 import _native as nv
    from optistate_b200.distributed import gather_columns
    from optistate_b200.synth import make_streams

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if a.dtype == "f64" else torch.float32
    esz = 8 if a.dtype == "f64" else 4
    n_local, T, S = a.traj_per_gpu, a.T, a.streams
    n_total = n_local * world
    first = rank * n_local

    # ---- synthetic inputs (host, pinned) -------------------------------------------------------------------
    st = make_streams(range(S), T)
    host = {k: torch.from_numpy(st[k]).to(dtype).pin_memory() for k in ("imu", "p", "dp", "contact", "f", "truth")}
    q_np, r_np = mc_noise(first, n_local, S)
    host["Q"] = torch.from_numpy(q_np).to(dtype).pin_memory()
    host["R"] = torch.from_numpy(r_np).to(dtype).pin_memory()
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    # label stream 2: the nominal member of every stream (u = v = 0), one small launch, untimed
    nominal = kf_batch(d["imu"], d["p"], d["dp"], d["contact"], d["f"], dtype=dtype, outputs=("x_steps",)).x_steps
    host["nominal"] = nominal.cpu().pin_memory()
    d["nominal"] = nominal
    out = {"summary": torch.empty((nv.SUMMARY_ROWS, n_local), dtype=dtype, device=dev),
           "status": torch.zeros(n_local, dtype=torch.int32, device=dev),
           "workspace": torch.empty(T * 10 * S * esz + 4 * S + 4096, dtype=torch.uint8, device=dev)}
    summary_host = torch.empty((nv.SUMMARY_ROWS, n_local), dtype=dtype).pin_memory()

    # N > 1: every GPU ends each step holding the summaries of ALL trajectories.  Default: the filter kernel stores them
    # into every GPU's copy itself (peer.PeerSummary); --gather nccl (or no peer access on this box): NCCL all-gather
    peer, gather_how = None, "single GPU"
    if world > 1:
        gather_how = "NCCL all-gather after the kernel"
        if a.gather == "fused":
            from optistate_b200.peer import PeerSummary
            try:
                peer = PeerSummary(n_total, dtype)
                gather_how = "fused into the filter kernel (NVLink peer stores) + one 4-byte NCCL all-reduce as barrier"
            except RuntimeError as e:  # agreed on by all ranks inside PeerSummary
                gather_how += f" ({e})"

    def step_nccl():
        res = kf_batch(d["imu"], d["p"], d["dp"], d["contact"], d["f"], Q=d["Q"], R=d["R"], n_traj=n_local, dtype=dtype,
                       stream_offset=first, truth=d["truth"], nominal=d["nominal"], outputs=("summary",), out=out,
                       q_kind=nv.MAT_DIAG_PER, r_kind=nv.MAT_DIAG_PER)
        if world > 1:
            return gather_columns(res.summary, n_total)
        return res.summary

    def step_fused():
        peer.wait()  # every rank is done with the previous step's gathered array
        kf_batch(d["imu"], d["p"], d["dp"], d["contact"], d["f"], Q=d["Q"], R=d["R"], n_traj=n_local, dtype=dtype,
                 stream_offset=first, truth=d["truth"], nominal=d["nominal"], outputs=("summary",), out=out,
                 q_kind=nv.MAT_DIAG_PER, r_kind=nv.MAT_DIAG_PER, summary_peers=peer)
        peer.wait()  # all kernels have finished: peer.tensor holds all n_total columns on every GPU
        return peer.tensor

    step_resident = step_fused if peer is not None else step_nccl

    # end-to-end through the public host-buffer API: every step uploads ALL of its inputs from pinned host memory and
    # downloads its summaries; the double-buffered pipeline overlaps step k+1's upload and step k-1's download with step
    # k's kernel, and the host reads each step's result one step later
    pipe = None
    e2e_state = {"prev": None, "sink": 0.0}

    def step_e2e():
        ticket = pipe.submit(host)
        if e2e_state["prev"] is not None:
            e2e_state["sink"] += float(pipe.result(e2e_state["prev"])[48, 0])  # the previous step's result, read on the host
        e2e_state["prev"] = ticket

    def finish_e2e():
        pipe.drain()
        e2e_state["sink"] += float(pipe.result(e2e_state["prev"])[48, 0])
        e2e_state["prev"] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = nv.ext().launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()  # work queued on other streams must be inside the timed region
            torch.cuda.synchronize()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), nv.ext().launch_count() - l0

    # ---- measured FMA peaks (roofline denominators) --------------------------------------------------------
    peak_flops, _ = fma_peak(dtype, 1 << 18)
    for _ in range(a.warmup):
        step_resident()
    with ClockSampler(local) as clk:
        ms, launches = timed(step_resident, a.steps)
    steps_total = n_total * T * a.steps
    value = steps_total / (ms * 1e-3)
    status_bad = int((out["status"] != 0).sum().item())
    gather_info = None
    if world > 1:
        gather_info = {"how": gather_how}
        if peer is not None:
            # the same job with the NCCL all-gather, and a check that the two gathered arrays are identical
            fused_copy = peer.tensor.clone()
            for _ in range(2):
                ref = step_nccl()
            gather_info["identical_to_nccl_gather"] = bool(torch.equal(fused_copy, ref))
            ms_n, _ = timed(step_nccl, a.steps)
            gather_info["nccl_gather_value"] = steps_total / (ms_n * 1e-3)
            gather_info["nccl_gather_ms_per_step"] = ms_n / a.steps
            del fused_copy, ref

    # secondary figures (rank 0, one GPU, outside the timed region; reported next to the headline, never part of it):
    # BASELINE configs[1] - 1,024 trajectories x 10,000 steps, FP64, per-step state dump (a latency case: 32 warps in
    # flight) - and the general JOINT path (reference operand order, four lanes per trajectory) on 131,072 trajectories
    secondary = None
    if rank == 0 and a.dtype == "f64":
        def once(fn):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = float("inf")
            for _ in range(3):
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e-3)
            return best
        reps = max(1, 10000 // T)  # the 1,000-step streams repeated in time: the data content does not matter for the rate
        st2 = {k: d[k][:, :, :1024].repeat(reps, 1, 1).contiguous() for k in ("imu", "p", "dp", "contact", "f")}
        t2 = once(lambda: kf_batch(st2["imu"], st2["p"], st2["dp"], st2["contact"], st2["f"], outputs=("x_steps", "p_trace", "k_gain")))
        tj = once(lambda: kf_batch(d["imu"][:200], d["p"][:200], d["dp"][:200], d["contact"][:200], d["f"][:200], n_traj=1 << 17,
                                   algo="joint", outputs=("x_final",)))
        n2, T2 = st2["imu"].shape[2], st2["imu"].shape[0]
        del st2
        from optistate_b200.mpc import mpc_forces
        from optistate_b200.synth import make_mpc_problems

        qp = [torch.from_numpy(v).to(dev) for v in make_mpc_problems(1 << 15)]
        tq = once(lambda: mpc_forces(*qp))
        trot = torch.where((torch.arange(1 << 15, device=dev) % 2 == 0)[None, :], torch.tensor([1.0, 0, 0, 1], device=dev, dtype=torch.float64)[:, None],
                           torch.tensor([0, 1.0, 1, 0], device=dev, dtype=torch.float64)[:, None]).contiguous()
        tqt = once(lambda: mpc_forces(qp[0], qp[1], qp[2], trot))
        secondary = {"cfg2_1024x10k_f64_steps_per_s": n2 * T2 / t2, "cfg2_seconds": t2,
                     "joint_path_f64_steps_per_s": (1 << 17) * 200 / tj, "force_mpc_qps_per_s": (1 << 15) / tq,
                     "force_mpc_trot_qps_per_s": (1 << 15) / tqt}

    e2e = None
    if not a.no_e2e:
        from optistate_b200.pipeline import KfHostPipeline

        pipe = KfHostPipeline(n_local, T, S, dtype=dtype, labels=("truth", "nominal"), stream_offset=first)
        for _ in range(max(2, a.warmup - 1)):
            step_e2e()
        finish_e2e()
        ms_e, _ = timed(step_e2e, a.steps, finish=finish_e2e)
        h2d = sum(host[k].numel() * host[k].element_size() for k in ("imu", "p", "dp", "contact", "f", "truth", "nominal", "Q", "R"))
        e2e = {"value": steps_total / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": summary_host.numel() * summary_host.element_size() * world, "ms_per_step": ms_e / a.steps,
               "api": "optistate_b200.pipeline.KfHostPipeline (kf_batch on pinned host buffers, double-buffered)"}

    if peer is not None:
        peer.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = clk.summary()
    per_gpu = value / world
    peak_tf = peak_flops / 1e12
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    lanes = 64 if a.dtype == "f64" else 128
    theoretical_tf = 148 * lanes * 2 * sm_max * 1e6 / 1e12
    achieved_tf = per_gpu * FLOPS_ALGORITHMIC / 1e12
    executed_tf = per_gpu * FLOPS_EXECUTED[a.dtype] / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
        "data": "synthetic", "config": workload_config(a, world), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {
            "bound": "fma", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
            "traffic": TRAFFIC_BYTES.get((a.dtype, a.traj_per_gpu, a.T, a.streams)), "algorithmic_bytes": algorithmic_bytes(a, esz),
            "note": f"per GPU; achieved = {int(FLOPS_ALGORITHMIC)} algorithmic flops/trajectory-step x steps/s; peak = {a.dtype} FMA issue peak "
                    "measured in this run (optistate_fma_peak); the kernel executes fewer flops than the algorithmic count "
                    "(sequential scalar updates on a packed symmetric P), hence frac can exceed 1 - see executed_*",
            "executed_flops_per_step": FLOPS_EXECUTED[a.dtype], "executed_tflops": executed_tf, "executed_frac_of_measured_peak": executed_tf / peak_tf,
            "theoretical_peak_tflops": theoretical_tf, "frac_of_theoretical": achieved_tf / theoretical_tf,
            "hbm_peak_gbs": _hbm_peak(),
        },
        "status_nonzero_trajectories": status_bad, "secondary": secondary,
    }
    if gather_info is not None:
        line["gather"] = gather_info
        line["config"]["sharding"] = f"contiguous blocks x{world}; summaries gathered on every GPU: {gather_how}"
    if world == 1 and not a.no_cpu_baseline:
        v, info = cpu_port_rate(a)
        line["cpu_baseline"] = dict({"value": v, "unit": UNIT}, **info)
        line["cpu_baseline"]["numpy_port_steps_per_s_1core"] = numpy_port_rate()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0  # B200_PROFILING.md fallback


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)
