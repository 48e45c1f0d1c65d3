"""Ad-hoc: where do the permutation elements of one synthetic sample go, by segment length class?"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, synth
ctx = genomic_b200.Context(0)
vals, off, lab, ids = synth.cohort([0], scale=1.0)
gp = Params(nperm=10000, rng_mode=RNG_MT19937_64, chain=False, seed=1, record_splits=True)
r = ctx.segment_batch(vals, off, gp, unit_ids=ids)
edges = [0, 4096, 16384, 32768, 65535, 1 << 30]
agg = collections.defaultdict(lambda: [0, 0, 0])
for s in r.splits:
    n = s["hi"] - s["lo"]
    k = next(i for i in range(len(edges) - 1) if edges[i] < n <= edges[i + 1])
    a = agg[k]; a[0] += 1; a[1] += s["perms_run"]; a[2] += s["perms_run"] * n
tot = sum(a[2] for a in agg.values())
for k in sorted(agg):
    a = agg[k]
    print(f"n in ({edges[k]},{edges[k+1]}]: splits {a[0]:4d} perms {a[1]:7d} perm_elems {a[2]:12d} ({100*a[2]/tot:.1f}%)")
big = sorted(r.splits, key=lambda s: -(s["perms_run"] * (s["hi"] - s["lo"])))[:12]
for s in big:
    print("unit", s["unit"], "n", s["hi"] - s["lo"], "perms", s["perms_run"], "exit", s["exit_code"], "ncpt", s["ncpt"])
