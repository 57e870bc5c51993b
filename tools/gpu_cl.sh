timeout 600 python -m pytest tests/test_mpc_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/closed_loop_rate.py 65536 30
timeout 300 python tools/closed_loop_rate.py 1024 100
