# lone-warp time profile of the decoupled-group FP64 kernel (1,024 trajectories = one warp per SM on 32 SMs)
mkdir -p gpurun_out/lone
export CASE_N=1024 CASE_T=400 CASE_S=1024
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kf_seq_tma --launch-skip 1 --launch-count 1 -f -o gpurun_out/lone/lone_xfinal python tools/ncu_case.py f64:tma:x_final > gpurun_out/lone/log1.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kf_seq_tma --launch-skip 1 --launch-count 1 -f -o gpurun_out/lone/lone_xsteps python tools/ncu_case.py f64:tma:x_steps > gpurun_out/lone/log2.txt 2>&1
tail -3 gpurun_out/lone/log1.txt gpurun_out/lone/log2.txt
python tools/cfg2_latency.py 2>&1 | grep sequential
ls -la gpurun_out/lone
