timeout 600 python -m pytest tests/test_mpc_gpu.py tests/test_reference_driver_gpu.py tests/test_features_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 120 python tools/closed_loop_rate.py 65536 20
timeout 120 python tools/closed_loop_rate.py 1024 200
timeout 120 python tools/closed_loop_rate.py 64 400
timeout 300 python examples/driver_pipeline.py 64 4063 --closed-loop 2>&1 | tail -1
