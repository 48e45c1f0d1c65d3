"""Per-kernel averages from an `ncu --set full` report -> markdown table + JSON (dram traffic per launch) under profiles/.
usage: python tools/ncu_kernel_summary.py <report.ncu-rep> <out_prefix>"""
import csv, io, json, subprocess, sys, collections
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def num(r, name):
    if name not in col: return None
    try: return float(r[col[name]].replace(",", ""))
    except ValueError: return None
def to_bytes(r, name):
    v = num(r, name)
    if v is None: return None
    u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
def to_ms(r, name):
    v = num(r, name)
    if v is None: return None
    u = units[col[name]].lower()
    return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(u, 1)
metrics = [("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
           ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
           ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
           ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           ("regs", "launch__registers_per_thread")]
agg = collections.OrderedDict()
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("cbsg::", "")
    a = agg.setdefault(name, collections.defaultdict(float))
    a["launches"] += 1
    a["ms"] += to_ms(r, "gpu__time_duration.sum") or 0.0
    a["dram_read"] += to_bytes(r, "dram__bytes_read.sum") or 0.0
    a["dram_write"] += to_bytes(r, "dram__bytes_write.sum") or 0.0
    a["inst"] += num(r, "smsp__inst_executed.sum") or 0.0
    w = to_ms(r, "gpu__time_duration.sum") or 0.0
    for key, m in metrics:
        v = num(r, m)
        if v is not None: a[key] += v * (w if key != "regs" else 1.0)
res = {}
lines = ["| kernel | launches | avg ms | dram read MB/launch | dram write MB/launch | achieved dram GB/s | issue % | warps % | fp64 % | alu % | lsu % | regs |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for name, a in agg.items():
    n, ms = a["launches"], a["ms"]
    w = lambda k: (a[k] / ms) if ms else 0.0
    res[name] = {"launches": int(n), "avg_ms": ms / n, "dram_read_bytes_per_launch": a["dram_read"] / n,
                 "dram_write_bytes_per_launch": a["dram_write"] / n, "warp_inst_per_launch": a["inst"] / n,
                 "issue_active_pct": w("issue_active_pct"), "warps_active_pct": w("warps_active_pct"),
                 "fp64_pipe_pct": w("fp64_pipe_pct"), "regs": a["regs"] / n}
    gbs = (a["dram_read"] + a["dram_write"]) / (ms * 1e-3) / 1e9 if ms else 0
    lines.append(f"| {name} | {int(n)} | {ms/n:.3f} | {a['dram_read']/n/1e6:.1f} | {a['dram_write']/n/1e6:.1f} | {gbs:.0f} | "
                 f"{w('issue_active_pct'):.1f} | {w('warps_active_pct'):.1f} | {w('fp64_pipe_pct'):.1f} | {w('alu_pipe_pct'):.1f} | "
                 f"{w('lsu_pipe_pct'):.1f} | {a['regs']/n:.0f} |")
open(out + ".md", "w").write(f"ncu --set full --clock-control none ({rep}); percentages are time-weighted averages over the captured launches\n\n" + "\n".join(lines) + "\n")
json.dump(res, open(out + ".json", "w"), indent=1)
print("\n".join(lines))
