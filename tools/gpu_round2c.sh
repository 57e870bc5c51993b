set -x
mkdir -p gpurun_out/r2i
nvidia-smi topo -m > gpurun_out/r2i/topo.txt 2>&1
timeout 300 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum --clock-control none -k regex:kf_seq_tma --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/r2i/nvlink_fused_gather.csv python tools/nvlink_probe.py > gpurun_out/r2i/nvlink_probe.log 2>&1
tail -3 gpurun_out/r2i/nvlink_probe.log
for n in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --traj-total 16777216 --steps 3 --warmup 3 > gpurun_out/r2i/bench_f64_16M_n$n.json 2> gpurun_out/r2i/bench_f64_16M_n$n.err
cut -c1-300 gpurun_out/r2i/bench_f64_16M_n$n.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 > gpurun_out/r2i/bench_f64_n8.json 2> gpurun_out/r2i/bench_f64_n8.err
cut -c1-300 gpurun_out/r2i/bench_f64_n8.json
