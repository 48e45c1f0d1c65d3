"""-m gpu: weighted CBS (cbs::segment_weighted, lib/cbs/CBS.cpp:1026-1099) on the CUDA path against the CPU oracle
(oracle/cbs_oracle.c, pinned to the compiled reference by tests/test_oracle.py) and, where oracle/_ref travelled to the
box, against the compiled reference itself.  MT replay: segment lengths, draws consumed and means are bit-identical."""
import numpy as np
import pytest

from helpers import f32, make_unit, pack
from oracle.pyoracle import SegParams
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, RNG_PHILOX

pytestmark = pytest.mark.gpu


def gparams(p: SegParams, **kw) -> Params:
    return Params(alpha=p.alpha, nperm=p.nperm, hybrid=p.hybrid, min_width=p.min_width, kmax=p.kmax, nmin=p.nmin,
                  eta=p.eta, tol=p.tol, do_smooth=False, rng_mode=RNG_PHILOX if p.rng_kind else RNG_MT19937_64,
                  chain=p.chain, seed=p.seed, undo_prune=p.undo_prune, undo_prune_cutoff=p.undo_prune_cutoff, **kw)


def make_weights(rng, n, kind):
    if kind == 0:
        return rng.uniform(0.5, 2.0, n)
    if kind == 1:  # few distinct values (ties in cw differences)
        return rng.choice([0.5, 1.0, 2.0], n)
    return np.ones(n)


CASE2_X = np.array([0.0] * 15 + [2.0] * 15 + [-1.5] * 15 + [0.0] * 15)
CASE2_W = np.array([1.0] * 15 + [0.5] * 15 + [2.0] * 15 + [1.0] * 15)


@pytest.mark.parametrize("alpha,nperm,hybrid,min_width", [(0.01, 200, False, 2), (0.05, 100, False, 3),
                                                          (0.01, 200, True, 2), (0.05, 100, True, 3)])
def test_weighted_kat_case2(ctx, alpha, nperm, hybrid, min_width):
    # tests/cbs_test.cpp:309-330 (inputs literal in tests/cbs_generate.R:91-92): 15/15/15/15, means 0/2/-1.5/0
    p = SegParams(alpha=alpha, nperm=nperm, hybrid=hybrid, min_width=min_width, seed=1, do_smooth=False)
    lengths, means, _ = ctx.segment_weighted(CASE2_X, CASE2_W, gparams(p))
    assert lengths.tolist() == [15, 15, 15, 15]
    assert np.allclose(means, [0.0, 2.0, -1.5, 0.0], atol=1e-9)


def test_weighted_matches_reference(ctx, ref):
    rng = np.random.default_rng(71)
    for trial in range(24):
        n = int(rng.integers(4, 2500))
        x = make_unit(rng, n, int(rng.integers(0, 5)))
        w = make_weights(rng, n, trial % 3)
        p = SegParams(nperm=int(rng.choice([50, 200, 1000])), alpha=float(rng.choice([0.01, 0.05])),
                      min_width=int(rng.choice([2, 3, 5])), do_smooth=False, seed=int(rng.integers(1, 100)),
                      undo_prune=bool(trial % 8 == 7))
        if n < 2 * p.min_width:
            continue
        eng = ref.rng(p.seed)
        wl, wm = ref.segment_weighted(x, w, p, eng)
        gl, gm, draws = ctx.segment_weighted(x, w, gparams(p, first_batch=int(rng.choice([16, 64, 256]))))
        assert np.array_equal(gl, wl), (trial, n, gl, wl)
        assert np.array_equal(gm, wm), (trial, n)
        assert ref.rng_equals(eng, p.seed, draws), (trial, n, draws)


def test_weighted_long_units_reference(ctx, ref):
    # one unit per shuffle size class, including > 65535 markers; null data so that the permutation loops run long
    rng = np.random.default_rng(72)
    for n in (6000, 30000, 70000):
        x = f32(rng.normal(0, 0.2, n))
        x[n // 3:] += 0.04
        w = rng.uniform(0.5, 2.0, n)
        p = SegParams(nperm=40, alpha=0.05, do_smooth=False, seed=4)
        eng = ref.rng(p.seed)
        wl, wm = ref.segment_weighted(x, w, p, eng)
        gl, gm, draws = ctx.segment_weighted(x, w, gparams(p, first_batch=24, max_batch=64))
        assert np.array_equal(gl, wl), (n, gl, wl)
        assert np.array_equal(gm, wm)
        assert ref.rng_equals(eng, p.seed, draws)


@pytest.mark.parametrize("mode", ["mt_unit", "mt_chain", "philox"])
def test_weighted_batch_matches_oracle(ctx, oracle, mode):
    rng = np.random.default_rng({"mt_unit": 73, "mt_chain": 74, "philox": 75}[mode])
    for trial in range(6):
        units = [make_unit(rng, int(rng.integers(1, 1500)), int(rng.integers(0, 5))) for _ in range(int(rng.integers(1, 7)))]
        if trial % 3 == 0:
            units.insert(1, np.zeros(0))
        vals, off = pack(units)
        w = make_weights(rng, len(vals), trial % 3)
        p = SegParams(nperm=int(rng.choice([50, 200, 1000])), alpha=float(rng.choice([0.01, 0.05])),
                      min_width=int(rng.choice([2, 3])), do_smooth=False, rng_kind=1 if mode == "philox" else 0,
                      chain=(mode == "mt_chain"), seed=int(rng.integers(1, 100)))
        want = oracle.segment_weighted_units(vals, w, off, p)
        got = ctx.segment_weighted_batch(vals, w, off, gparams(p, first_batch=int(rng.choice([16, 64, 256]))))
        assert np.array_equal(got.seg_count, want["seg_count"]), trial
        assert np.array_equal(got.lengths, want["lengths"]), trial
        assert np.array_equal(got.means, want["means"]), trial
        if p.rng_kind == 0:
            assert np.array_equal(got.draws, want["draws"]), trial


def test_weighted_rejects(ctx):
    x = np.zeros(10)
    with pytest.raises(ValueError):  # invalid argument
        ctx.segment_weighted(x, np.r_[np.ones(9), 0.0], Params(do_smooth=False))  # non-positive weight
    with pytest.raises(ValueError):
        ctx.segment_weighted(x, np.ones(9), Params(do_smooth=False))  # size mismatch


def test_weighted_hybrid_matches_reference(ctx, ref):
    # hybrid p-values: delta from the weights (getmncwt), tail probability, hwtmaxp permutations (CBS.cpp:908-921)
    rng = np.random.default_rng(76)
    for trial in range(12):
        n = int(rng.integers(201, 3000))
        x = make_unit(rng, n, int(rng.integers(0, 4)))
        w = make_weights(rng, n, trial % 3)
        p = SegParams(nperm=int(rng.choice([50, 200, 1000])), alpha=float(rng.choice([0.01, 0.05])),
                      min_width=int(rng.choice([2, 3, 5])), do_smooth=False, seed=int(rng.integers(1, 100)),
                      hybrid=True, kmax=int(rng.choice([25, 10])), nmin=200)
        eng = ref.rng(p.seed)
        wl, wm = ref.segment_weighted(x, w, p, eng)
        gl, gm, draws = ctx.segment_weighted(x, w, gparams(p, first_batch=int(rng.choice([16, 64, 256]))))
        assert np.array_equal(gl, wl), (trial, n, gl, wl)
        assert np.allclose(gm, wm, rtol=1e-9, atol=0), (trial, n)
        assert ref.rng_equals(eng, p.seed, draws), (trial, n, draws)


@pytest.mark.parametrize("mode", ["mt_unit", "philox"])
def test_weighted_hybrid_batch_matches_oracle(ctx, oracle, mode):
    rng = np.random.default_rng({"mt_unit": 77, "philox": 78}[mode])
    units = [make_unit(rng, n, 1) for n in (150, 400, 2500, 9000)]
    vals, off = pack(units)
    w = make_weights(rng, len(vals), 0)
    p = SegParams(nperm=300, alpha=0.01, min_width=2, do_smooth=False, rng_kind=1 if mode == "philox" else 0, chain=False,
                  seed=5, hybrid=True)
    want = oracle.segment_weighted_units(vals, w, off, p)
    got = ctx.segment_weighted_batch(vals, w, off, gparams(p, first_batch=64))
    assert np.array_equal(got.seg_count, want["seg_count"])
    assert np.array_equal(got.lengths, want["lengths"])
    assert np.allclose(got.means, want["means"], rtol=1e-9, atol=0)
    if p.rng_kind == 0:
        assert np.array_equal(got.draws, want["draws"])


def test_weighted_full_size_sample_properties(ctx, oracle):
    """weighted CBS on one SNP6-scale sample (1.8 M markers, nperm 10 000): lengths partition every unit, means are the
    sequential weighted sums, the call is deterministic, splitting it into two calls changes nothing, and the two smallest
    chromosomes equal the oracle bit for bit (draws included)."""
    from genomic_b200 import synth
    vals, off, lab, ids = synth.cohort([0], scale=1.0)
    x = vals.astype(np.float64)
    w = np.random.default_rng(20260101).uniform(0.5, 2.0, len(x))
    gp = Params(nperm=10000, alpha=0.01, do_smooth=False, rng_mode=RNG_MT19937_64, chain=False, seed=1)
    a = ctx.segment_weighted_batch(x, w, off, gp, unit_ids=ids)
    b = ctx.segment_weighted_batch(x, w, off, gp, unit_ids=ids)
    assert np.array_equal(a.lengths, b.lengths) and np.array_equal(a.means, b.means) and np.array_equal(a.draws, b.draws)
    for u in range(len(off) - 1):
        s0, s1 = int(a.seg_offsets[u]), int(a.seg_offsets[u + 1])
        lens = a.lengths[s0:s1]
        assert lens.min() >= 1 and int(lens.sum()) == int(off[u + 1] - off[u])
        pos = int(off[u])
        for k, ln in enumerate(lens):
            sw = np.cumsum(w[pos:pos + ln])[-1]
            swx = np.cumsum(w[pos:pos + ln] * x[pos:pos + ln])[-1]  # sequential, like CBS.cpp:1093-1095
            assert a.means[s0 + k] == swx / sw, (u, k)
            pos += int(ln)
    cutu = 10
    c1 = ctx.segment_weighted_batch(x[:off[cutu]], w[:off[cutu]], off[:cutu + 1], gp, unit_ids=ids[:cutu])
    c2 = ctx.segment_weighted_batch(x[off[cutu]:], w[off[cutu]:], off[cutu:] - off[cutu], gp, unit_ids=ids[cutu:])
    assert np.array_equal(np.concatenate([c1.lengths, c2.lengths]), a.lengths)
    assert np.array_equal(np.concatenate([c1.means, c2.means]), a.means)
    assert np.array_equal(np.concatenate([c1.draws, c2.draws]), a.draws)
    p = SegParams(nperm=10000, alpha=0.01, do_smooth=False, seed=1)
    for u in np.argsort(np.diff(off))[:2]:
        u = int(u)
        rng = oracle.rng_mt(1)
        wl, wm = oracle.segment_weighted(x[off[u]:off[u + 1]], w[off[u]:off[u + 1]], p, rng)
        s0, s1 = int(a.seg_offsets[u]), int(a.seg_offsets[u + 1])
        assert np.array_equal(a.lengths[s0:s1], wl) and np.array_equal(a.means[s0:s1], wm)
        assert int(a.draws[u]) == int(rng.draws)


def test_weighted_edge_inputs(ctx, ref):
    # constant data (all-equal shortcut, CBS.cpp:1051), units shorter than 2*min_width, n = 1, wide-ranging weights
    rng = np.random.default_rng(79)
    p = SegParams(nperm=200, alpha=0.05, min_width=3, do_smooth=False, seed=3)
    cases = [(np.full(40, 0.25), rng.uniform(0.5, 2.0, 40)), (np.array([0.1]), np.array([1.0])),
             (np.array([0.1, 0.2, 0.3, 0.4, 0.5]), np.ones(5)), (np.r_[np.zeros(30), np.ones(30)], np.ones(60))]
    for k in range(6):
        n = int(rng.integers(60, 900))
        x = make_unit(rng, n, k % 4)
        cases.append((x, 10.0 ** rng.uniform(-3, 3, n)))
    for i, (x, w) in enumerate(cases):
        eng = ref.rng(p.seed)
        wl, wm = ref.segment_weighted(x, w, p, eng)
        gl, gm, draws = ctx.segment_weighted(x, w, gparams(p, first_batch=32))
        assert np.array_equal(gl, wl), (i, gl, wl)
        assert np.array_equal(gm, wm), i
        assert ref.rng_equals(eng, p.seed, draws), i
