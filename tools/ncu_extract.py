"""Pull the judged metrics out of an .ncu-rep (all captured launches): python tools/ncu_extract.py <rep> > <json>"""
import csv, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio"]
res = []
for r in rows[2:]:
    d = {}
    for w in WANT:
        if w in h:
            i = h.index(w)
            d[w] = r[i] + ((" " + units[i]) if units[i] and w != "Kernel Name" else "")
    res.append(d)
print(json.dumps({"report": rep, "launches": res}, indent=1))
