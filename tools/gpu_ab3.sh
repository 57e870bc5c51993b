export CASE_T=1000
echo "== base"; LD_LIBRARY_PATH=variants/base python tools/rate.py f64:summary f32:summary
echo "== tree"; python tools/cfg2_latency.py 2>&1 | grep sequential; python tools/rate.py f64:summary f32:summary
