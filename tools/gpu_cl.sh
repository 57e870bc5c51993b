timeout 600 python -m pytest tests/test_mpc_gpu.py -m gpu -q -x 2>&1 | tail -12
