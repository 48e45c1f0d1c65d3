import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import genomic_b200
from oracle.pyoracle import Ref
ctx = genomic_b200.Context(0); ref = Ref()
def centred(x):
    cur = x - np.cumsum(x)[-1] / len(x); return cur, float(np.cumsum(cur * cur)[-1])
bad = 0; tot = 0
for n in (400, 900, 1600, 2500, 10000):
    for pat in ("alt", "saw4", "saw10", "blocks", "zero_mean_steps"):
        if pat == "alt": x = np.tile([1.0, -1.0], n // 2)
        elif pat == "saw4": x = np.tile([1.0, 1.0, -1.0, -1.0], n // 4)
        elif pat == "saw10": x = np.tile([1.0]*5 + [-1.0]*5, n // 10)
        elif pat == "blocks":
            b = int(round(np.sqrt(n))); x = np.tile([1.0]*(b//2) + [-1.0]*(b - b//2), n // b + 1)[:n]
        else: x = np.tile([2.0, -1.0, -1.0], n // 3 + 1)[:n]
        for raw in (True, False):
            xc, tss = (x, float((x*x).sum())) if raw else centred(x)
            for al0 in (2, 3):
                g = ctx.tmaxo(xc, tss, al0); w = ref.tmaxo(xc, tss, al0); tot += 1
                if g != w: bad += 1; print("UNW MISMATCH", n, pat, raw, al0, g, w)
                wts = np.ones(n) if al0 == 2 else np.tile([1.0, 2.0], n // 2 + 1)[:n]
                g = ctx.wtmaxo(xc, wts, tss, al0); w = ref.wtmaxo(xc, wts, tss, al0); tot += 1
                if g != w: bad += 1; print("W MISMATCH", n, pat, raw, al0, g, w)
print("cases", tot, "mismatches", bad)
