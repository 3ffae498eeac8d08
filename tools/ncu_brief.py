"""print the handful of ncu raw-page metrics that decide where a memory-bound kernel loses time: python tools/ncu_brief.py x.ncu-rep [row]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct"]
for row in r[2:]:
    print("-" * 60)
    for i, x in enumerate(h):
        if x in want or ("average_warps_issue_stalled" in x and float(row[i] or 0) > 0.3):
            print("%-80s %-10s %s" % (x, r[1][i], row[i]))
