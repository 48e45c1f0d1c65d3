"""Per-source-line instruction / stall-sample shares from `ncu -i X --page source --csv --print-source cuda,sass`."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
want_file = sys.argv[2] if len(sys.argv) > 2 else "kernels.cuh"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
fname = None; hdr = None
names = ("stall_long_sb", "stall_wait", "stall_short_sb", "stall_barrier", "stall_math", "stall_not_selected", "stall_mio",
         "stall_lg", "stall_branch_resolving", "stall_dispatch", "stall_selected")
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; iex = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples"); idx = {n: hdr.index(n) for n in names if n in hdr}; continue
    if hdr is None or fname is None or not fname.endswith(want_file) or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    a = agg[ln]
    a[0] += int(r[iex] or 0); a[1] += int(r[ismp] or 0); a[2] = r[1]
    for n, i in idx.items(): a[3][n] += int(r[i] or 0)
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
print("file", want_file, "warp instructions", tot, "samples", tots)
for ln, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    t2 = ", ".join(f"{k[6:]}={v}" for k, v in a[3].most_common(2))
    print(f"{ln:5d} inst {100*a[0]/tot:5.1f}% smp {100*a[1]/tots:5.1f}% [{t2}] {a[2].strip()[:90]}")
