"""Ad-hoc: per-round kernel timeline of one synthetic sample (CBS_GPU_DEBUG_ROUNDS), after two warm-up calls."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, RNG_PHILOX, synth
nsamp = int(os.environ.get("NSAMP", "1"))
mode = os.environ.get("MODE", "mt")
ctx = genomic_b200.Context(0)
vals, off, lab, ids = synth.cohort(list(range(nsamp)), scale=float(os.environ.get("SCALE", "1.0")))
gp = Params(nperm=10000, rng_mode=RNG_PHILOX if mode == "philox" else RNG_MT19937_64, chain=False, seed=1)
for rep in range(4):
    if rep == 3:
        os.environ["CBS_GPU_DEBUG_ROUNDS"] = "1"
        ctx.set_profiling(events=True)
    t0 = time.time()
    r = ctx.segment_batch(vals, off, gp, unit_ids=ids)
    print("rep", rep, "wall %.1f ms" % (1e3 * (time.time() - t0)), "rounds", r.rounds, "ms", r.ms, flush=True)
print({k: round(v, 2) for k, v in ctx.last_kernel_ms().items()})
