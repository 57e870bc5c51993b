timeout 600 python -m pytest tests/test_mpc_gpu.py tests/test_reference_driver_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "== tree"; timeout 300 python tools/closed_loop_rate.py 65536 30; timeout 300 python tools/mpc_rate.py 131072 trot
echo "== base"; LD_LIBRARY_PATH=variants/base timeout 300 python tools/closed_loop_rate.py 65536 30
echo "== tree"; timeout 300 python tools/closed_loop_rate.py 65536 30
./tools/probes/fp64_issue_mix
