"""Summarise an `ncu --set full` capture of the Mimi decoder's tcgen05 launches (tools/mimi_once.py) into a text table:
    python tools/ncu_mimi_summary.py gpurun_out/<name>.ncu-rep > profiles/<tag>_mimi_kernels_ncu_summary.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum"]
idx = {w: hdr.index(w) for w in want if w in hdr}
kn = hdr.index("Kernel Name")
print("ncu --set full --clock-control none --import-source on, capture %s (python tools/mimi_once.py 144 100: chunks of 36 utterances x 100 frames)" % sys.argv[1])
for i, r in enumerate(rows[2:]):
    print("== launch %d: %s" % (i, r[kn][:90]))
    for w, j in idx.items():
        print("   %s = %s %s" % (w, r[j], units[j]))
