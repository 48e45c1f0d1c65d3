"""Ad-hoc performance probe (not the bench): per-kernel times of one synthetic sample."""
import os, sys, time, faulthandler, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
faulthandler.dump_traceback_later(int(os.environ.get("PROBE_WATCHDOG", "600")), exit=True)
import numpy as np
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, RNG_PHILOX, synth

scales = [float(s) for s in os.environ.get("SCALES", "0.1,1.0").split(",")]
modes = os.environ.get("MODES", "mt_unit,philox").split(",")
nperm = int(os.environ.get("NPERM", "10000"))
nsamp = int(os.environ.get("NSAMP", "1"))
ctx = genomic_b200.Context(0)
print("fp64 pipe: %.3f T inst/s" % ctx.measure_fp64(), flush=True)
for scale in scales:
    vals, off, lab, ids = synth.cohort(list(range(nsamp)), scale=scale)
    for mode in modes:
        gp = Params(nperm=nperm, rng_mode=RNG_PHILOX if mode == "philox" else RNG_MT19937_64, chain=False, seed=1,
                    first_batch=int(os.environ.get("FIRST_BATCH", "0")), max_batch=int(os.environ.get("MAX_BATCH", "0")))
        for rep in range(3):
            ctx.set_profiling(events=(rep == 1), counters=(rep == 2))
            t0 = time.time()
            r = ctx.segment_batch(vals, off, gp, unit_ids=ids)
            dt = time.time() - t0
            line = dict(scale=scale, mode=mode, rep=rep, markers=int(off[-1]), wall_s=round(dt, 4), rounds=r.rounds,
                        perms=r.perms_run, launches=r.kernel_launches, segs=int(len(r.lengths)),
                        ms={k: round(v, 3) for k, v in r.ms.items()}, mps=round(off[-1] / dt))
            if rep == 1:
                line["kernel_ms"] = {k: round(v, 3) for k, v in ctx.last_kernel_ms().items()}
            if rep == 2:
                a, s = ctx.last_arc_evals()
                line["arcs"] = a; line["slots"] = s
            print(json.dumps(line), flush=True)
print("done", flush=True)
