"""Ad-hoc: step time of the bench workload against the permutation batch sizes (first_batch, max_batch), configurations
interleaved so that box-level drift affects all of them alike."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, synth
ctx = genomic_b200.Context(0)
vals, off, lab, ids = synth.cohort([0], scale=1.0)
cfgs = [(256, 4096), (256, 10000), (192, 4096), (384, 4096), (128, 4096), (256, 2048)]
ts = {c: [] for c in cfgs}
ref = None
for rep in range(10):
    for c in cfgs:
        gp = Params(nperm=10000, rng_mode=RNG_MT19937_64, chain=False, seed=1, first_batch=c[0], max_batch=c[1])
        t0 = time.perf_counter()
        r = ctx.segment_batch(vals, off, gp, unit_ids=ids)
        ts[c].append(1e3 * (time.perf_counter() - t0))
        if ref is None:
            ref = r
        assert np.array_equal(r.lengths, ref.lengths) and np.array_equal(r.means, ref.means) and np.array_equal(r.draws, ref.draws)
for c in cfgs:
    v = sorted(ts[c][1:])
    print(f"first_batch={c[0]:5d} max_batch={c[1]:6d} median={v[len(v)//2]:7.1f} ms  p25={v[len(v)//4]:7.1f}  min={v[0]:7.1f}  mean={sum(v)/len(v):7.1f}", flush=True)
