"""CPU tests of the host logic: synthetic generator contract, trajectory sharding + the summary all-gather over
gloo with world_size 2, the bench's CPU arm and its Monte-Carlo noise, argument classification of kf_batch."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_generator_contract():
    from optistate_b200.synth import SEED0_CHECK, make_stream, make_streams, monte_carlo_noise

    s = make_stream(0, 3)  # SURVEY 8(d) check values were taken at T = 3
    assert s["p"][0, 0] == SEED0_CHECK["p[0,0]"] and s["dp"][0, 0] == SEED0_CHECK["dp[0,0]"]
    assert s["imu"][0, 0] == pytest.approx(SEED0_CHECK["imu[0,0]"], rel=1e-15) and s["f"][0, 2] == pytest.approx(SEED0_CHECK["f[0,2]"], rel=1e-15)
    s = make_stream(3, 200)
    assert set(np.unique(s["contact"])) == {0.0, 1.0} and (s["contact"].sum(axis=1) == 2).all()  # trot: never all-swing
    assert (s["contact"][:25, [0, 3]] == 1).all() and (s["contact"][25:50, [1, 2]] == 1).all()
    assert (s["f"][:, [0, 1, 3, 4, 6, 7, 9, 10]] == 0).all() and (s["f"][:, 2][s["contact"][:, 0] == 0] == 0).all()
    st = make_streams([3, 4], 200)
    assert st["imu"].shape == (200, 6, 2) and np.array_equal(st["p"][:, :, 0], s["p"])
    q, r = monte_carlo_noise(np.arange(6), np.full(12, 0.01), np.full(10, 0.01), nominal_every=2)
    assert (q[:, :2] == 0.01).all() and (q[:, 2:] != 0.01).all() and (q > 0.01 * 10**-0.5).all() and (q < 0.01 * 10**0.5).all()


def test_shard_range_covers_everything_once():
    from optistate_b200.distributed import shard_range, shard_sizes

    for n, w in ((16, 4), (17, 4), (3, 8), (1 << 24, 8)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert sum(shard_sizes(n, w)) == n and max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from optistate_b200.distributed import gather_columns, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
n_total = 11  # uneven shards: 6 + 5
b, e = shard_range(n_total, 2, dist.get_rank())
local = torch.arange(b, e, dtype=torch.float64).repeat(52, 1) + 1000.0 * torch.arange(52, dtype=torch.float64)[:, None]
full = gather_columns(local, n_total)
want = torch.arange(n_total, dtype=torch.float64).repeat(52, 1) + 1000.0 * torch.arange(52, dtype=torch.float64)[:, None]
assert full.shape == (52, n_total) and torch.equal(full, want), full
dist.barrier(); dist.destroy_process_group(); print("ok")
"""


def test_summary_all_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, err[-2000:]


_SHARE_WORKER = """
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from optistate_b200.distributed import stream_exchange_group
from optistate_b200.pipeline import stream_share_plan
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
grp = stream_exchange_group()  # gloo: a plain group (the high-priority option is NCCL's)
rank = dist.get_rank()
sizes = {"imu": 7 * 6 * 3, "p": 7 * 12 * 3, "contact": 7 * 4 * 3 + 1}  # odd total: the flat allocation is padded
rng = np.random.default_rng(5)
host = {k: torch.from_numpy(rng.standard_normal(n)) for k, n in sizes.items()}  # the same on both ranks
flat_len, ranges, pieces = stream_share_plan(sizes, 2, rank)
flat = torch.full((flat_len,), float("nan"), dtype=torch.float64)
for k, src0, dst0, cnt in pieces:  # this rank's share only
    flat[dst0:dst0 + cnt] = host[k][src0:src0 + cnt]
share = flat_len // 2
mine = flat[rank * share:(rank + 1) * share].clone()
parts = [torch.empty(share, dtype=torch.float64) for _ in range(2)]
dist.all_gather(parts, mine, group=grp)  # (gloo has no all_gather_into_tensor: same exchange, list form)
full = torch.cat(parts)
for k, (a, b) in ranges.items():
    assert torch.equal(full[a:b], host[k]), k
dist.barrier(); dist.destroy_process_group(); print("ok")
"""


def test_shared_stream_upload_plan_tiles_every_array_once_and_exchanges_over_gloo(tmp_path):
    """KfHostPipeline's shared upload: every rank uploads its 1/world share of the flat stream allocation, one all-gather completes
    it on every rank.  The plan is pure arithmetic; the exchange is checked with world_size 2 over gloo."""
    from optistate_b200.pipeline import stream_share_plan

    sizes = {"a": 10, "b": 3, "c": 24, "d": 1}
    for world in (1, 2, 3, 4, 8):
        seen = {k: np.zeros(n, dtype=np.int64) for k, n in sizes.items()}
        for rank in range(world):
            flat_len, ranges, pieces = stream_share_plan(sizes, world, rank)
            assert flat_len % world == 0 and flat_len >= sum(sizes.values()) and ranges["a"] == (0, 10) and ranges["d"] == (37, 38)
            share = flat_len // world
            for k, src0, dst0, cnt in pieces:
                assert dst0 == ranges[k][0] + src0 and rank * share <= dst0 and dst0 + cnt <= (rank + 1) * share
                seen[k][src0:src0 + cnt] += 1
        assert all((v == 1).all() for v in seen.values()), world
    script = tmp_path / "w.py"
    script.write_text(_SHARE_WORKER)
    port = str(27500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, err[-2000:]


def test_bench_noise_is_a_function_of_the_member_id():
    sys.path.insert(0, ROOT)
    import bench

    qa, ra = bench.mc_noise(0, 70000, 1024)
    qb, rb = bench.mc_noise(65000, 5000, 1024)
    assert np.array_equal(qa[:, 65000:], qb) and np.array_equal(ra[:, 65000:], rb)  # independent of the shard boundary
    assert (qa[:, :1024] == bench.Q_DIAG[:, None]).all() and not (qa[:, 1024:2048] == bench.Q_DIAG[:, None]).any()


def test_bench_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--T", "50"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "trajectory-steps/s" and line["value"] > 1e3
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-side argument handling only")
def test_noise_argument_classification_needs_no_gpu():
    from optistate_b200 import _native as nv
    from optistate_b200.batch import _noise

    cpu = torch.device("cpu")
    t, k = _noise(np.diag(np.full(12, 0.01)), 12, 5, "Q", torch.float64, cpu)
    assert k == nv.MAT_DIAG and t.shape == (12,)  # np.diag(...) as the reference builds Q: recognised as diagonal
    dense = np.diag(np.full(12, 0.01)); dense[0, 1] = 1e-3
    assert _noise(dense, 12, 5, "Q", torch.float64, cpu)[1] == nv.MAT_DENSE
    assert _noise(np.ones((12, 5)), 12, 5, "Q", torch.float64, cpu)[1] == nv.MAT_DIAG_PER
    assert _noise(np.ones((144, 5)), 12, 5, "P0", torch.float64, cpu)[1] == nv.MAT_DENSE_PER
    with pytest.raises(ValueError):
        _noise(np.ones((7, 3)), 12, 5, "Q", torch.float64, cpu)


def test_gru_pipeline_restatement_shapes_and_reference_model_compatibility():
    """The consumer restated in oracle/gru_pipeline.py has the reference RNN's parameter layout; when the reference tree is
    present (build container) the two classes produce identical outputs from the same state_dict."""
    from oracle import gru_pipeline, ref_shim

    rows = np.random.default_rng(0).standard_normal((50, 60))
    norm, lo, hi = gru_pipeline.min_max_normalise(rows)
    assert norm.min() == 0.0 and norm.max() == 1.0
    w = gru_pipeline.windows(norm, gru_pipeline.seeded_latent(50))
    assert w.shape == (41, 10, 188) and w.dtype == np.float32 and np.array_equal(w[3, 2, :60], norm[5].astype(np.float32))
    model = gru_pipeline.seeded_consumer()
    assert sum(p.numel() for p in model.parameters()) == 422424  # gru_model.RNN(188,128,4,24)
    if ref_shim.available():
        sys.path.insert(0, ref_shim.REF_ROOT)
        from gru.gru_model import RNN

        ref_model = RNN(188, 128, 4, 24, torch.device("cpu")).eval()
        ref_model.load_state_dict(model.state_dict())
        x = torch.from_numpy(w)
        with torch.no_grad():
            assert torch.equal(ref_model(x), model(x))


def test_identification_restatement_matches_reference_class_including_z_aliasing():
    """oracle/identify_numpy.py vs the same loop driven through the unmodified reference class (build container only)."""
    from oracle import identify_numpy, ref_shim
    from optistate_b200.synth import make_stream

    if not ref_shim.available():
        pytest.skip("reference tree only exists in the build container")
    s = make_stream(77, 120)
    rng = np.random.default_rng(2)
    gt = s["truth"] + 0.02 * rng.standard_normal(s["truth"].shape)
    q_ref, r_ref = identify_numpy.identify_with_reference_class(gt, s["imu"], s["p"], s["dp"], s["contact"], s["f"])
    q, r = identify_numpy.identify(gt, s["imu"], s["p"], s["dp"], s["contact"], s["f"], alias_last_measurement=True)
    assert np.abs(q / q_ref - 1).max() < 1e-12 and np.abs(r / r_ref - 1).max() < 1e-12
    q2, r2 = identify_numpy.identify(gt, s["imu"], s["p"], s["dp"], s["contact"], s["f"], alias_last_measurement=False)
    assert np.allclose(q2, q) and not np.allclose(r2, r)  # the aliasing only changes R


@pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the pre-launch argument checks only")
def test_kf_batch_rejects_inconsistent_arguments_before_touching_the_device():
    """Shape errors are raised by the host layer; without a GPU the first thing kf_batch does is refuse to run."""
    from optistate_b200 import kf_batch
    from optistate_b200.batch import _stream_tensor

    cpu = torch.device("cpu")
    with pytest.raises(ValueError, match=r"\[T, 6, S\]"):
        _stream_tensor(np.zeros((5, 7, 2)), 6, "imu", torch.float64, cpu)
    t, T, S = _stream_tensor(np.zeros((5, 6)), 6, "imu", torch.float64, cpu)  # a single stream may omit the last axis
    assert (T, S) == (5, 1) and t.shape == (5, 6, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kf_batch(np.zeros((5, 6, 1)), np.zeros((5, 12, 1)), np.zeros((5, 12, 1)), np.ones((5, 4, 1)), np.zeros((5, 12, 1)))
