set -x
mkdir -p gpurun_out/mpc
timeout 900 python -m pytest tests/test_mpc_gpu.py -m gpu -q -x > gpurun_out/mpc/pytest_mpc.txt 2>&1; echo "rc=$?" >> gpurun_out/mpc/pytest_mpc.txt
tail -25 gpurun_out/mpc/pytest_mpc.txt
timeout 300 python tools/mpc_rate.py 32768 trot > gpurun_out/mpc/rate_trot.txt 2>&1; cat gpurun_out/mpc/rate_trot.txt
timeout 300 python tools/mpc_rate.py 32768 > gpurun_out/mpc/rate_all.txt 2>&1; cat gpurun_out/mpc/rate_all.txt
timeout 300 python tools/closed_loop_rate.py 65536 30 > gpurun_out/mpc/closed_loop.txt 2>&1; cat gpurun_out/mpc/closed_loop.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_mpc_gi --launch-skip 1 --launch-count 1 -o gpurun_out/mpc/mpc_gi python tools/mpc_rate.py 32768 trot > gpurun_out/mpc/ncu.log 2>&1
