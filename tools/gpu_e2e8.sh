mkdir -p gpurun_out/e2e
for mode in shared replicated; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-secondary --e2e-upload $mode > gpurun_out/e2e/n8_$mode.json 2> gpurun_out/e2e/n8_$mode.err
done
for mode in shared replicated; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --no-secondary --e2e-upload $mode > gpurun_out/e2e/n4_$mode.json 2> gpurun_out/e2e/n4_$mode.err
done
python - <<'PY'
import json
for f in ("n8_shared","n8_replicated","n4_shared","n4_replicated"):
    try:
        d=json.loads(open(f"gpurun_out/e2e/{f}.json").read().strip().splitlines()[-1]); print(f, "%.4g" % d["value"], "%.4g" % d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"])
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/e2e/{f}.err").read()[-1500:])
PY
