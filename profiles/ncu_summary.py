"""Key numbers of one `ncu --set full` report.  usage: python profiles/ncu_summary.py <rep>"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "Kernel Name"]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in WANT:
        for h in hdr:
            if h == k:
                print("%-70s %-12s %s" % (h, units[hdr.index(h)], d[h]))
    st = sorted(((float(d[h]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h in hdr
                 if h.startswith(STALL) and h.endswith("per_issue_active.ratio") and d[h]), reverse=True)
    print("stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in st[:8]))
