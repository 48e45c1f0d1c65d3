"""Summary of an `ncu --set full` capture of ONE round of the worklist -> markdown table + profiles/ncu_traffic.json.

usage: python tools/ncu_kernel_summary.py <report.ncu-rep> <out_prefix> <round> <markers_x_permutations of that round> [capture note]

The capture is taken with tools/one_step.py (exactly one call, so launch k of a per-round kernel belongs to round k) and
`-k regex:"k_chain|k_shuffle|k_scan" -s <9*round> -c 9`; the round's planned markers x permutations come from the same
command's plain run with CBS_GPU_DEBUG_ROUNDS=1 ("[rounds] planned perms / markers per round").  k_chain and k_scan process
every element of the round in their one launch; the shuffle launches of a round (one per segment-length class) are added up."""
import collections
import csv
import io
import json
import subprocess
import sys

rep, out, rnd, elements = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
note = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return None


def scaled(r, name, table):
    v = num(r, name)
    return None if v is None else v * table.get(units[col[name]].lower(), 1)


BYTES = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
MS = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3}
METRICS = [("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
           ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           ("alu pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
           ("lsu pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
           ("l1/smem %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
           ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
           ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           ("regs", "launch__registers_per_thread")]


def group_of(name):
    if "k_shuffle" in name or "k_perm" in name:
        return "k_shuffle"
    if "k_chain" in name:
        return "k_chain"
    if "k_scan" in name:
        return "k_scan"
    return None


lines = ["| launch | ms | dram read MB | dram write MB | " + " | ".join(m for m, _ in METRICS) + " |",
         "|---|---|---|---|" + "---|" * len(METRICS)]
agg = collections.OrderedDict()
for r in data:
    name = r[col["Kernel Name"]].replace("cbsg::", "").replace("void ", "")
    short = name.split("(")[0]
    ms = scaled(r, "gpu__time_duration.sum", MS) or 0.0
    rd = scaled(r, "dram__bytes_read.sum", BYTES) or 0.0
    wr = scaled(r, "dram__bytes_write.sum", BYTES) or 0.0
    lines.append(f"| {short} | {ms:.3f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | " +
                 " | ".join("-" if num(r, m) is None else f"{num(r, m):.1f}" for _, m in METRICS) + " |")
    g = group_of(short)
    if g:
        a = agg.setdefault(g, dict(ms=0.0, dram_read_bytes=0.0, dram_write_bytes=0.0, launches=0, warp_inst=0.0))
        a["ms"] += ms; a["dram_read_bytes"] += rd; a["dram_write_bytes"] += wr; a["launches"] += 1
        a["warp_inst"] += num(r, "smsp__inst_executed.sum") or 0.0
res = {"_capture": f"ncu --set full --clock-control none, round {rnd} of one synthetic SNP6-scale sample, {elements:.4g} markers x "
                   f"permutations planned in that round" + (f"; {note}" if note else "")}
lines += ["", f"Round {rnd}: {elements:.4g} markers x permutations.", "",
          "| kernel (all launches of the round) | ms | dram bytes / element | algorithmic bytes / element | ratio | GB/s (dram) | warp instructions / element |",
          "|---|---|---|---|---|---|---|"]
ALG = {"k_shuffle": 16.0, "k_chain": 16.0, "k_scan": 8.0}
for g, a in agg.items():
    a["elements"] = elements
    res[g] = a
    per = (a["dram_read_bytes"] + a["dram_write_bytes"]) / elements
    lines.append(f"| {g} | {a['ms']:.3f} | {per:.2f} | {ALG[g]:.0f} | {per / ALG[g]:.2f} | "
                 f"{(a['dram_read_bytes'] + a['dram_write_bytes']) / (a['ms'] * 1e-3) / 1e9:.0f} | {a['warp_inst'] / elements:.2f} |")
open(out + ".md", "w").write(f"`ncu --set full --clock-control none` ({rep}), one row per captured launch; cold-cache, serialised launches: "
                             f"compare shares and per-element traffic, not absolute times.\n\n" + "\n".join(lines) + "\n")
json.dump(res, open(out + ".json", "w"), indent=1)
print("\n".join(lines))
