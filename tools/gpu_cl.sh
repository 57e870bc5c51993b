timeout 600 python -m pytest tests/test_mpc_gpu.py -m gpu -q 2>&1 | tail -15
timeout 300 python tools/closed_loop_rate.py 65536 30
WARM_ROUNDS=-1 timeout 300 python tools/closed_loop_rate.py 65536 30
timeout 300 python tools/mpc_rate.py 32768 trot
