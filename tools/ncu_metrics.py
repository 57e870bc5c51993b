"""Developer helper: the handful of ncu raw-page metrics the profiles/*_metrics.txt summaries hold, one `name [unit] value`
line each (the format bench.kernel_profile parses).
usage: python tools/ncu_metrics.py report.ncu-rep [kernel-name-regex] [launch-index] [extra-metric ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "nvltx__bytes.sum", "nvlrx__bytes.sum", "lts__t_bytes.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else ""
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
want = WANT + sys.argv[4:]
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"] + (["--kernel-name", "regex:" + rx] if rx else [])
rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
hdr, units, data = rows[0], rows[1], rows[2:]
r = data[which]
ix = {h: i for i, h in enumerate(hdr)}
print("Kernel Name []", r[ix["Kernel Name"]])
for m in want:
    if m in ix and r[ix[m]] != "":
        print(f"{m} [{units[ix[m]]}] {r[ix[m]]}")
