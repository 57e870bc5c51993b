set -x
mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest_gpu.txt
tail -6 gpurun_out/r2d/pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2d/bench_f64.json 2> gpurun_out/r2d/bench_f64.err; cut -c1-400 gpurun_out/r2d/bench_f64.json
timeout 300 python bench.py --dtype f32 --no-secondary --no-cpu-baseline > gpurun_out/r2d/bench_f32.json 2> gpurun_out/r2d/bench_f32.err; cut -c1-200 gpurun_out/r2d/bench_f32.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_seq_tma --launch-skip 2 --launch-count 1 -o gpurun_out/r2d/f64_blk python bench.py --steps 1 --warmup 1 --no-e2e --no-secondary --no-cpu-baseline > gpurun_out/r2d/ncu_f64.log 2>&1
