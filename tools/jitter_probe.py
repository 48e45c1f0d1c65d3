"""Ad-hoc: per-step wall time next to per-kernel event sums, to tell GPU-side from host-side jitter."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, synth
ctx = genomic_b200.Context(0)
vals, off, lab, ids = synth.cohort([0], scale=1.0)
gp = Params(nperm=10000, rng_mode=RNG_MT19937_64, chain=False, seed=1)
for prof in (False, True):
    ctx.set_profiling(events=prof)
    for rep in range(14):
        t0 = time.perf_counter()
        r = ctx.segment_batch(vals, off, gp, unit_ids=ids)
        dt = 1e3 * (time.perf_counter() - t0)
        line = f"prof={int(prof)} rep={rep:2d} wall={dt:7.1f} ms seg={r.ms['segment']:7.1f} smooth={r.ms['smooth']:5.1f} h2d={r.ms['h2d']:5.1f}"
        if prof:
            k = ctx.last_kernel_ms()
            line += " | " + " ".join(f"{n}={v:.1f}" for n, v in k.items() if v > 1.0)
        print(line, flush=True)
