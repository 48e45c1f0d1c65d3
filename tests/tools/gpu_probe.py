"""Ad-hoc GPU probe: times a few calls per mode with a watchdog so a hang costs seconds, not minutes."""
import os, sys, time, faulthandler
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
faulthandler.dump_traceback_later(int(os.environ.get("PROBE_WATCHDOG", "240")), exit=True)
import numpy as np
from helpers import make_unit, pack
from oracle.pyoracle import Oracle, SegParams
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, RNG_PHILOX
O = Oracle()
ctx = genomic_b200.Context(0)
rng = np.random.default_rng(3)
for mode in ("mt_chain", "mt_unit", "philox"):
    for trial in range(3):
        units = [make_unit(rng, int(rng.integers(1, 1500)), int(rng.integers(0, 5))) for _ in range(4)]
        vals, off = pack(units)
        p = SegParams(nperm=1000, alpha=0.01, do_smooth=False, rng_kind=1 if mode == "philox" else 0, chain=(mode == "mt_chain"), seed=5)
        t0 = time.time(); want = O.segment_units(vals, off, np.ones(len(off) - 1, np.int32), p); t1 = time.time()
        gp = Params(alpha=p.alpha, nperm=p.nperm, do_smooth=False, rng_mode=RNG_PHILOX if p.rng_kind else RNG_MT19937_64, chain=p.chain, seed=p.seed)
        print(mode, trial, "oracle %.3fs" % (t1 - t0), "units", [len(u) for u in units], flush=True)
        got = ctx.segment_batch(vals, off, gp); t2 = time.time()
        ok = np.array_equal(got.seg_count, want["seg_count"]) and np.array_equal(got.lengths, want["lengths"]) and np.array_equal(got.means, want["means"])
        print("   gpu %.3fs rounds=%d perms=%d launches=%d ms=%s ok=%s" % (t2 - t1, got.rounds, got.perms_run, got.kernel_launches, got.ms, ok), flush=True)
print("probe done", flush=True)
