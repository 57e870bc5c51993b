mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2k/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k/pytest_gpu.txt
tail -3 gpurun_out/r2k/pytest_gpu.txt
(timeout 120 python tools/mpc_rate.py 32768 trot; timeout 120 python tools/mpc_rate.py 32768; timeout 120 python tools/mpc_rate.py 16384 stand; timeout 120 python tools/mpc_rate.py 16384 three; timeout 120 python tools/closed_loop_rate.py 65536 20; timeout 120 python tools/closed_loop_rate.py 65536 30) > gpurun_out/r2k/mpc_rates.txt 2>&1
cat gpurun_out/r2k/mpc_rates.txt | grep -v "per step"
