timeout 600 python -m pytest tests/test_mpc_gpu.py tests/test_reference_driver_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 120 python tools/mpc_rate.py 32768 trot
timeout 120 python tools/mpc_rate.py 32768
timeout 120 python tools/mpc_rate.py 16384 stand
timeout 120 python tools/closed_loop_rate.py 65536 20 | tail -1
timeout 120 python tools/closed_loop_rate.py 65536 30 | tail -1
