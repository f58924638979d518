"""Summarise an ncu --set full report: `python profiles/summarize_ncu.py <file.ncu-rep> > profiles/<name>_summary.txt`.
Prints one line per metric of interest with the values of every captured launch (read with `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_sleeping", "smsp__pcsamp_warps_issue_stalled_selected"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
print(f"# ncu --set full --clock-control none --import-source on (cold cache, serialised launches): {len(data)} captured launches")
for k in KEEP:
    for i, h in enumerate(hdr):
        if h == k or (k.startswith("smsp__pcsamp") and h == k):
            vals = [r[i] for r in data]
            if k == "Kernel Name":
                vals = [v.split("(")[0] for v in vals]
            print(f"{h} [{units[i]}] " + " | ".join(vals))
