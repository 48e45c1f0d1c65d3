"""Ad-hoc: step time of the bench workload against the permutation batch sizes (first_batch, max_batch)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, synth
ctx = genomic_b200.Context(0)
vals, off, lab, ids = synth.cohort([0], scale=1.0)
ref = None
for fb, mb in ((256, 4096), (128, 4096), (512, 4096), (1024, 4096), (256, 8192), (512, 8192), (384, 4096), (256, 2048)):
    gp = Params(nperm=10000, rng_mode=RNG_MT19937_64, chain=False, seed=1, first_batch=fb, max_batch=mb)
    ts = []
    for rep in range(7):
        t0 = time.perf_counter()
        r = ctx.segment_batch(vals, off, gp, unit_ids=ids)
        ts.append(1e3 * (time.perf_counter() - t0))
    if ref is None:
        ref = r
    same = np.array_equal(r.lengths, ref.lengths) and np.array_equal(r.means, ref.means) and np.array_equal(r.draws, ref.draws)
    ts.sort()
    print(f"first_batch={fb:5d} max_batch={mb:5d} median={ts[3]:7.1f} ms min={ts[0]:7.1f} rounds={r.rounds} perms={r.perms_run} elems={r.perm_elems} same={same}", flush=True)
