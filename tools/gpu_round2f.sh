mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest_gpu.txt
tail -4 gpurun_out/r2f/pytest_gpu.txt
timeout 120 python tools/dropin_rate.py
