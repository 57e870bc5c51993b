set -x
mkdir -p gpurun_out/r2h
nvidia-smi topo -m > gpurun_out/r2h/topo.txt 2>&1
ncu --query-metrics 2>/dev/null | grep -i nvl > gpurun_out/r2h/nvl_metrics_available.txt
timeout 600 python -m pytest tests/test_peer_gpu.py tests/test_parity_gpu.py -m gpu -q -k "peer or fused or config3_shape or odd_row or stored" > gpurun_out/r2h/pytest_peer.txt 2>&1; echo "rc=$?" >> gpurun_out/r2h/pytest_peer.txt
timeout 600 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum --clock-control none -k regex:kf_seq_tma --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/r2h/nvlink_fused_gather.csv python tools/nvlink_probe.py > gpurun_out/r2h/nvlink_probe.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --traj-total 16777216 --steps 3 --warmup 3 > gpurun_out/r2h/bench_f64_16M_n2.json 2> gpurun_out/r2h/bench_f64_16M_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2h/bench_f64_n2.json 2> gpurun_out/r2h/bench_f64_n2.err
tail -3 gpurun_out/r2h/pytest_peer.txt; cat gpurun_out/r2h/nvlink_probe.log | tail -3; cut -c1-600 gpurun_out/r2h/bench_f64_16M_n2.json
