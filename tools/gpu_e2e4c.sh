mkdir -p gpurun_out/e2e
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 4 --no-secondary --no-cpu-baseline --e2e-upload shared > gpurun_out/e2e/n4_shared_v2.json 2> gpurun_out/e2e/n4_shared_v2.err
python - <<'PY'
import json
for f in ("n4_shared_v2",):
    try:
        d=json.loads(open(f"gpurun_out/e2e/{f}.json").read().strip().splitlines()[-1]); print(f, "%.4g" % d["value"], "%.4g" % d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"].get("identical_to_resident_path"))
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/e2e/{f}.err").read()[-2500:])
PY
