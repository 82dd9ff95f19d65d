"""Key metrics of an .ncu-rep (ncu -i ... --page raw --csv) as a short text table."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for vals in rows[2:]:
    for i, h in enumerate(hdr):
        if h in keys or ("stalled" in h and "per_issue_active.ratio" in h):
            print("%-86s %-16s %s" % (h, units[i], vals[i]))
    print()
