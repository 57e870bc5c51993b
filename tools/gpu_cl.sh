timeout 600 python -m pytest tests/test_mpc_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 120 python tools/mpc_rate.py 32768 trot
timeout 120 python tools/mpc_rate.py 32768
timeout 120 python tools/mpc_rate.py 16384 stand
timeout 120 python tools/mpc_rate.py 16384 three
