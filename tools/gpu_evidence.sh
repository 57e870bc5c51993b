# Round-2 evidence on one B200: tests, bench lines, ncu captures of the dominant kernels, launch list, smoke.  Outputs: gpurun_out/r2g/
set -x
O=gpurun_out/r2g
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_f64.json 2> $O/bench_f64.err
timeout 300 python bench.py --dtype f32 --no-secondary --no-cpu-baseline > $O/bench_f32.json 2> $O/bench_f32.err
timeout 300 python bench.py --structure full --no-secondary --no-cpu-baseline > $O/bench_f64_full.json 2> $O/bench_f64_full.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_seq_tma --launch-skip 2 --launch-count 1 -o $O/f64_blk python bench.py --steps 1 --warmup 1 --no-e2e --no-secondary --no-cpu-baseline > $O/ncu_f64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_seq_tma --launch-skip 2 --launch-count 1 -o $O/f32_blk python bench.py --dtype f32 --steps 1 --warmup 1 --no-e2e --no-secondary --no-cpu-baseline > $O/ncu_f32.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kf_mpc_gi --launch-skip 1 --launch-count 1 -o $O/mpc_gi python tools/mpc_rate.py 32768 trot > $O/ncu_mpc.log 2>&1
timeout 300 ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum --clock-control none -k regex:fma_peak --launch-count 2 --csv --log-file $O/fma_peak_ncu.csv python -c "
import torch
from optistate_b200 import fma_peak
print(fma_peak(torch.float64, 1 << 21))" > $O/fma_peak.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1
timeout 120 python tools/mpc_rate.py 32768 trot > $O/mpc_rate_trot.txt 2>&1
timeout 120 python tools/mpc_rate.py 32768 > $O/mpc_rate_all.txt 2>&1
timeout 120 python tools/closed_loop_rate.py 65536 30 > $O/closed_loop.txt 2>&1
cut -c1-300 $O/bench_f64.json; cat $O/smoke.txt | tail -3
