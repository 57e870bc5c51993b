export CASE_T=1000
echo "== tree"; python tools/cfg2_latency.py 2>&1 | grep sequential; python tools/rate.py f64:summary f32:summary
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
