mkdir -p gpurun_out/r2j
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2j/smoke.txt 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/r2j/smoke.txt
timeout 600 python bench.py > gpurun_out/r2j/bench_f64.json 2> gpurun_out/r2j/bench_f64.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2j/bench_f64.json
