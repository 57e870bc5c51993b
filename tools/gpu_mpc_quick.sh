mkdir -p gpurun_out/mpc
timeout 600 python -m pytest tests/test_mpc_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/mpc_rate.py 32768 trot
timeout 300 python tools/mpc_rate.py 131072 trot
timeout 300 python tools/closed_loop_rate.py 65536 30
