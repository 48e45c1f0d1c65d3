"""ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`) -> markdown table per kernel: launches, total ms, share.
usage: python tools/launch_list_summary.py <launches.csv> <out.md> "<one-line description of the command>" """
import collections
import csv
import re
import sys

src, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(open(src, errors="replace")) if r and not r[0].startswith("==")]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
kn, mv, mu = col["Kernel Name"], col["Metric Value"], col["Metric Unit"]
scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn].replace("cbsg::", "").replace("void ", ""))
    name = re.sub(r"^cub::.*?(Device\w+Kernel|\w+Kernel).*", r"cub \1", name)
    try:
        ms = float(r[mv].replace(",", "")) * scale.get(r[mu], 1e-6)
    except ValueError:
        continue
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(v[1] for v in tot.values())
lines = [note, "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {name} | {n} | {ms:.3f} | {100 * ms / total:.1f}% |")
lines.append(f"| all | {sum(v[0] for v in tot.values())} | {total:.3f} | 100% |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:24]))
