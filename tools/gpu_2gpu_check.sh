mkdir -p gpurun_out/n2
timeout 600 python -m pytest tests/test_peer_gpu.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/n2/pytest_peer.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --no-secondary > gpurun_out/n2/bench_n2.json 2> gpurun_out/n2/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/n2/ref_n2.json 2> gpurun_out/n2/ref_n2.err
cut -c1-400 gpurun_out/n2/bench_n2.json; tail -3 gpurun_out/n2/bench_n2.err; cut -c1-300 gpurun_out/n2/ref_n2.json
