"""Summarise an .ncu-rep (ncu --set full) into a small CSV + markdown table under profiles/."""
import csv, subprocess, sys, io, os
rep, out_prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["ID", "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
with open(out_prefix + ".csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([a for a, _ in idx]); w.writerow([units[i] for _, i in idx])
    for r in data: w.writerow([r[i] for _, i in idx])
with open(out_prefix + ".md", "w") as f:
    f.write("| " + " | ".join(a for a, _ in idx) + " |\n|" + "---|" * len(idx) + "\n")
    for r in data: f.write("| " + " | ".join(r[i] for _, i in idx) + " |\n")
print("wrote", out_prefix + ".csv/.md", len(data), "launches")
