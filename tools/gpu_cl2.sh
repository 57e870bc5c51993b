mkdir -p gpurun_out/mpc
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/mpc/closed_loop_launches.csv python tools/closed_loop_rate.py 65536 8 > gpurun_out/mpc/cl_ncu.log 2>&1
tail -2 gpurun_out/mpc/cl_ncu.log
