mkdir -p gpurun_out/dropin
timeout 600 python -m pytest tests/test_dropin_gpu.py tests/test_abi_ctypes_gpu.py tests/test_reference_driver_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/dropin_rate.py --profile 2>&1 | tail -40
