import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle.pyoracle import Oracle, SegParams
from helpers import make_unit, pack
import genomic_b200
from test_gpu_weighted import gparams, make_weights
o=Oracle(); ctx=genomic_b200.Context(0)
rng=np.random.default_rng(73)
for trial in range(6):
    units=[make_unit(rng,int(rng.integers(1,1500)),int(rng.integers(0,5))) for _ in range(int(rng.integers(1,7)))]
    if trial%3==0: units.insert(1,np.zeros(0))
    vals,off=pack(units)
    w=make_weights(rng,len(vals),trial%3)
    p=SegParams(nperm=int(rng.choice([50,200,1000])),alpha=float(rng.choice([0.01,0.05])),min_width=int(rng.choice([2,3])),do_smooth=False,rng_kind=0,chain=False,seed=int(rng.integers(1,100)))
    fb=int(rng.choice([16,64,256]))
    want=o.segment_weighted_units(vals,w,off,p)
    gp=gparams(p,first_batch=fb); gp.record_splits=True
    got=ctx.segment_weighted_batch(vals,w,off,gp)
    print(trial, np.array_equal(got.lengths,want["lengths"]), np.array_equal(got.draws,want["draws"]), got.draws.tolist(), want["draws"].tolist(), np.diff(off).tolist(), p.nperm, p.alpha, fb)
    if not np.array_equal(got.draws,want["draws"]):
        for sp in got.splits: print(sp)
