timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_peer_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-secondary --no-e2e | cut -c1-160
