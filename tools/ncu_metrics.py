"""Print the headline metrics of every kernel in an ncu report: python tools/ncu_metrics.py X.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for w in want:
    if w not in hdr:
        continue
    i = hdr.index(w)
    vals = [r[i][:26] for r in data]
    if w.startswith("smsp__average_warps_issue_stalled_") and all(float(v or 0) < 0.3 for v in vals):
        continue
    print(w.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")[:74].ljust(76), units[i][:10].ljust(10), " | ".join(vals))
