mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/final/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee gpurun_out/final/smoke.txt
timeout 300 python tools/dropin_rate.py 2>&1 | tail -4 | tee gpurun_out/final/dropin.txt
