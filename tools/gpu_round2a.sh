set -x
mkdir -p gpurun_out/r2a
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2a/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r2a/bench_f64.json 2> gpurun_out/r2a/bench_f64.err; echo "rc=$?"
timeout 300 python bench.py --structure full --no-secondary --no-cpu-baseline > gpurun_out/r2a/bench_f64_full.json 2> gpurun_out/r2a/bench_f64_full.err
timeout 300 python bench.py --dtype f32 --no-secondary --no-cpu-baseline > gpurun_out/r2a/bench_f32.json 2> gpurun_out/r2a/bench_f32.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_seq_tma --launch-skip 2 --launch-count 1 -o gpurun_out/r2a/f64_blk python bench.py --steps 1 --warmup 1 --no-e2e --no-secondary --no-cpu-baseline > gpurun_out/r2a/ncu_f64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_seq_tma --launch-skip 2 --launch-count 1 -o gpurun_out/r2a/f32_blk python bench.py --dtype f32 --steps 1 --warmup 1 --no-e2e --no-secondary --no-cpu-baseline > gpurun_out/r2a/ncu_f32.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2a/launches.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r2a/smoke.txt 2>&1
tail -5 gpurun_out/r2a/pytest_gpu.txt
cat gpurun_out/r2a/bench_f64.json | cut -c1-1500
