mkdir -p gpurun_out/r2e
timeout 600 python -m pytest tests/test_reference_driver_gpu.py tests/test_dropin_gpu.py tests/test_mpc_gpu.py -m gpu -q -x > gpurun_out/r2e/pytest.txt 2>&1
tail -30 gpurun_out/r2e/pytest.txt
timeout 120 python tools/dropin_rate.py
