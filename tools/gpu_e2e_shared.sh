mkdir -p gpurun_out/e2e
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --no-secondary --no-cpu-baseline --e2e-upload shared --e2e-trace > gpurun_out/e2e/n8_shared_v2.json 2> gpurun_out/e2e/n8_shared_v2.err
python - <<'PY'
import json
for f in ("n8_shared_v2",):
    try:
        d=json.loads(open(f"gpurun_out/e2e/{f}.json").read().strip().splitlines()[-1]); print(f, "%.4g" % d["value"], "%.4g" % d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"].get("identical_to_resident_path"))
        for row in d["e2e"]["timeline_ms"]: print(row)
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/e2e/{f}.err").read()[-2500:])
PY
