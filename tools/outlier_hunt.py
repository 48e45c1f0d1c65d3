"""Ad-hoc: many calls on one synthetic sample with the event timeline on; keeps the timelines of the slowest and of a typical call."""
import os, sys, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, synth
ctx = genomic_b200.Context(0)
vals, off, lab, ids = synth.cohort([0])
gp = Params(nperm=10000, rng_mode=RNG_MT19937_64, chain=False, seed=1)
n = int(os.environ.get("CALLS", "60"))
prof = os.environ.get("PROF", "1") == "1"
for _ in range(3):
    ctx.segment_batch(vals, off, gp, unit_ids=ids)
if prof:
    os.environ["CBS_GPU_DEBUG_ROUNDS"] = "1"
    ctx.set_profiling(events=True)
walls = []
for k in range(n):
    sys.stderr.write("\n=== call %d\n" % k); sys.stderr.flush()
    t0 = time.time()
    r = ctx.segment_batch(vals, off, gp, unit_ids=ids)
    walls.append((1e3 * (time.time() - t0), dict(r.ms)))
med = sorted(w for w, _ in walls)[n // 2]
print("median wall %.1f ms" % med)
for k, (w, ms) in enumerate(walls):
    if w > 1.08 * med:
        print("call", k, "wall %.1f" % w, {a: round(b, 1) for a, b in ms.items()})
