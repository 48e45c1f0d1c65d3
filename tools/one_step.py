"""Ad-hoc: exactly one segment_batch call on one synthetic sample (for ncu captures: launch k of a kernel = round k)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, RNG_PHILOX, synth
ctx = genomic_b200.Context(0)
vals, off, lab, ids = synth.cohort(list(range(int(os.environ.get("NSAMP", "1")))), scale=float(os.environ.get("SCALE", "1.0")))
gp = Params(nperm=10000, rng_mode=RNG_PHILOX if os.environ.get("MODE", "mt") == "philox" else RNG_MT19937_64, chain=False, seed=1)
r = ctx.segment_batch(vals, off, gp, unit_ids=ids)
print("rounds", r.rounds, "perms", r.perms_run, "perm_elems", r.perm_elems, "launches", r.kernel_launches, "ms", r.ms)
