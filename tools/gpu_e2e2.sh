mkdir -p gpurun_out/e2e
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "host_pipeline" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-secondary > gpurun_out/e2e/n2_shared.json 2> gpurun_out/e2e/n2_shared.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --no-secondary --e2e-upload replicated > gpurun_out/e2e/n2_repl.json 2> gpurun_out/e2e/n2_repl.err
python - <<'PY'
import json
for f in ("n2_shared","n2_repl"):
    try:
        d=json.loads(open(f"gpurun_out/e2e/{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["e2e"])
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/e2e/{f}.err").read()[-1500:])
PY
