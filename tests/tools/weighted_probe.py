"""Weighted CBS (cbs::segment_weighted) on one synthetic SNP6-scale sample: wall time per call, per-kernel event times,
and a parity check of the two smallest chromosomes against the CPU oracle.  Prints one JSON line.

    python tests/tools/weighted_probe.py [--scale 1.0] [--reps 3] [--nperm 10000] [--check]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import genomic_b200  # noqa: E402
from genomic_b200 import Params, RNG_MT19937_64, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--nperm", type=int, default=10000)
    ap.add_argument("--no-warmup", action="store_true", help="skip the small warm-up call (ncu: the first k_wscan<1> launch is then round 1 of the sample)")
    ap.add_argument("--check", action="store_true", help="compare the two smallest chromosomes with the CPU oracle (slow at scale 1)")
    args = ap.parse_args()
    values, off, lab, ids = synth.cohort([0], scale=args.scale)
    x = values.astype(np.float64)
    rng = np.random.default_rng(20260101)
    w = rng.uniform(0.5, 2.0, len(x))  # per-marker weights (DNAcopy: inverse variances)
    p = Params(alpha=0.01, nperm=args.nperm, do_smooth=False, rng_mode=RNG_MT19937_64, chain=False, seed=1)
    ctx = genomic_b200.Context(0)
    if not args.no_warmup:
        ctx.segment_weighted_batch(x[: off[2]], w[: off[2]], off[:3], p)  # warm-up
    times = []
    res = None
    for _ in range(args.reps):
        t0 = time.perf_counter()
        res = ctx.segment_weighted_batch(x, w, off, p, unit_ids=ids)
        times.append((time.perf_counter() - t0) * 1e3)
    ctx.set_profiling(events=True)
    ctx.segment_weighted_batch(x, w, off, p, unit_ids=ids)
    kms = ctx.last_kernel_ms()
    ctx.set_profiling(events=False)
    out = {
        "workload": f"1 synthetic SNP6-scale sample, {len(x)} markers x 23 chromosomes, weighted CBS, nperm={args.nperm}, "
                    "MT replay, no smoothing",
        "ms_per_call": [round(t, 2) for t in times],
        "markers_samples_per_s": len(x) / (min(times) * 1e-3),
        "segments": int(res.lengths.size), "perms_run": res.perms_run, "perm_elements": res.perm_elems, "rounds": res.rounds,
        "kernel_ms": {k: round(v, 3) for k, v in kms.items()},
    }
    if args.check:
        from oracle.pyoracle import Oracle, SegParams, build
        build(ref=False)
        orc = Oracle()
        order = np.argsort(np.diff(off))[:2]
        ok = True
        t0 = time.perf_counter()
        for u in order:
            a, b = int(off[u]), int(off[u + 1])
            wl, wm = orc.segment_weighted(x[a:b], w[a:b], SegParams(alpha=0.01, nperm=args.nperm, do_smooth=False, seed=1))
            s0, s1 = int(res.seg_offsets[u]), int(res.seg_offsets[u + 1])
            ok = ok and np.array_equal(res.lengths[s0:s1], wl) and np.array_equal(res.means[s0:s1], wm)
        out["oracle_check"] = {"units": [int(u) for u in order], "bit_identical": bool(ok),
                               "cpu_s": round(time.perf_counter() - t0, 2),
                               "cpu_markers_per_s": float(sum(off[u + 1] - off[u] for u in order) / (time.perf_counter() - t0))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
