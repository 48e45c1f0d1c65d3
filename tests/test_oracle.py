"""CPU (-m "not gpu"): pin the oracle.  (1) the reference's own golden vectors / KATs for this
path; (2) bit-for-bit agreement of the C restatement with the UNMODIFIED reference
sources compiled into oracle/_ref (skipped where /root/reference was never available)."""
import os

import numpy as np
import pytest

from cnio import cohort_from_cn, is_log_scale, read_cn, seg_text
from helpers import f32, make_unit
from oracle.pyoracle import SegParams

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


# ---- (1) golden vectors of the reference ------------------------------------------------------
def test_golden_cli_seg_oracle(oracle):
    # tests/cna_test.cpp:258-270
    names, positions, _, values, off, lab, units = cohort_from_cn(os.path.join(GOLD, "segment_cli_case1_input.cn"))
    r = oracle.segment_units(values, off, lab, SegParams(chain=True))
    txt = seg_text(names, positions, units, r["seg_count"], r["lengths"], r["means"])
    assert txt == open(os.path.join(GOLD, "segment_cli_case1_expected.seg")).read()


def test_golden_cli_seg_ref(ref):
    names, positions, _, values, off, lab, units = cohort_from_cn(os.path.join(GOLD, "segment_cli_case1_input.cn"))
    r = ref.segment_units(values, off, lab, SegParams(chain=True))
    txt = seg_text(names, positions, units, r["seg_count"], r["lengths"], r["means"])
    assert txt == open(os.path.join(GOLD, "segment_cli_case1_expected.seg")).read()


def test_not_log_scale_rejected():
    # tests/cna_test.cpp:272-282 / src/cna_segment.hpp:109-125
    assert not is_log_scale(read_cn(os.path.join(GOLD, "segment_cli_not_logscale_input.cn"))[2])
    assert is_log_scale(read_cn(os.path.join(GOLD, "segment_cli_case1_input.cn"))[2])


def test_kat_case1_tmaxo(oracle):
    # tests/cbs_test.cpp:154-177, inputs literal in tests/cbs_generate.R:90-92
    x = np.array([0.0] * 20 + [1.5] * 20 + [0.0] * 20)
    tss = float((x * x).sum() - x.sum() ** 2 / len(x))
    stat, s, e = oracle.tmaxo(x, tss, 2, False)
    assert (s, e) == (0, 58) and stat > 1000.0
    stat, s, e = oracle.tmaxo(x, tss, 2, True)
    assert (s, e) == (0, 58) and stat > 10.0


@pytest.mark.parametrize("alpha,nperm,hybrid,min_width", [(0.01, 200, False, 2), (0.05, 100, False, 3),
                                                          (0.01, 200, True, 2), (0.05, 100, True, 3)])
def test_kat_case1_segment(oracle, alpha, nperm, hybrid, min_width):
    # tests/cbs_test.cpp:287-307: lengths 20/20/20, means 0/1.5/0 (<= 1e-9)
    x = np.array([0.0] * 20 + [1.5] * 20 + [0.0] * 20)
    lengths, means = oracle.segment(x, SegParams(alpha=alpha, nperm=nperm, hybrid=hybrid, min_width=min_width, seed=1))
    assert lengths.tolist() == [20, 20, 20]
    assert np.allclose(means, [0.0, 1.5, 0.0], atol=1e-9)


def test_smooth_identity_fixtures(oracle):
    # tests/smooth_generate.R:37-51 literal inputs: big outliers inflate the trimmed SD, output == input
    x1 = np.array([0.0, 0.1, -0.1, 0.05, 8.0, 0.0, -0.05, 0.1, 0.02, -0.02])
    out = oracle.smooth(x1, np.ones(len(x1), np.int32))
    assert out.shape == x1.shape
    x2 = x1.copy()
    x2[3] = np.nan
    out2 = oracle.smooth(x2, np.ones(len(x2), np.int32))
    assert np.isnan(out2[3])


# ---- (2) restatement == compiled reference ------------------------------------------------------
def test_rng_matches_reference(oracle, ref):
    import ctypes as C
    r, rr = oracle.rng_mt(1), ref.rng(1)
    for i in range(2000):
        assert oracle.lib.orc_rng_unif(C.byref(r)) == ref.lib.ref_rng_next_canonical(rr.h)


def test_tmaxo_matches_reference(oracle, ref):
    rng = np.random.default_rng(0)
    for trial in range(150):
        n = int(rng.integers(4, 3000))
        x = make_unit(rng, n, trial % 5)
        xc = x - x.mean()
        tss = float((xc * xc).sum())
        for al0 in (2, 3, 5):
            if n < 2 * al0:
                continue
            assert oracle.tmaxo(xc, tss, al0) == ref.tmaxo(xc, tss, al0)
            a, b = oracle.tmaxo(xc, tss, al0, True), ref.tmaxo(xc, tss, al0, True)
            assert a == b or (np.isnan(a[0]) and np.isnan(b[0]))


def test_perm_pieces_match_reference(oracle, ref):
    rng = np.random.default_rng(1)
    for trial in range(30):
        n = int(rng.integers(5, 2000))
        x = rng.normal(0, 0.2, n)
        x -= x.mean()
        a = oracle.xperm(x, oracle.rng_mt(7))
        b = ref.xperm(x, ref.rng(7))
        assert np.array_equal(a, b)
        n1 = int(rng.integers(1, n))
        ro, rr = oracle.rng_mt(3), ref.rng(3)
        assert oracle.tpermp(n1, n - n1, x, 100, ro) == ref.tpermp(n1, n - n1, x, 100, rr)
        assert ref.rng_equals(rr, 3, ro.draws)
        if n > 210:
            tss = float((x * x).sum())
            assert oracle.htmaxp(a, tss, 25) == ref.htmaxp(a, tss, 25)
    for b in (2.0, 3.5, 4.2, 5.0):
        for m in (300, 5000):
            assert oracle.tailp(b, 26 / m, m) == ref.tailp(b, 26 / m, m)


def test_segment_matches_reference(oracle, ref):
    rng = np.random.default_rng(2)
    for trial in range(60):
        n = int(rng.integers(4, 1200))
        x = make_unit(rng, n, trial % 5)
        for hybrid in (False, True):
            p = SegParams(nperm=int(rng.choice([100, 200, 1000])), alpha=float(rng.choice([0.01, 0.05])), hybrid=hybrid,
                          min_width=int(rng.choice([2, 3, 5])), undo_prune=bool(trial % 5 == 0))
            ro, rr = oracle.rng_mt(1), ref.rng(1)
            la, ma = oracle.segment(x, p, ro)
            lb, mb = ref.segment(x, p, rr)
            assert np.array_equal(la, lb) and np.array_equal(ma, mb)
            assert ref.rng_equals(rr, 1, ro.draws)


def test_smooth_matches_reference(oracle, ref):
    rng = np.random.default_rng(3)
    fired = 0
    for trial in range(60):
        n = int(rng.integers(0, 3000))
        x = rng.normal(0, 0.2, n)
        for i in (rng.integers(0, max(n, 1), max(1, n // 50)) if n else []):
            x[i] += rng.choice([-1, 1]) * (3 + abs(rng.normal()))
        if trial % 3 == 0 and n:
            for i in rng.integers(0, n, 3):
                x[i] = rng.choice([np.nan, np.inf, -np.inf])
        chrom = np.sort(rng.integers(1, 4, n)).astype(np.int32) if trial % 2 else np.ones(n, np.int32)
        reg = int(rng.choice([0, 1, 2, 10]))
        trim = float(rng.choice([0.025, 0.01, 0.1, 0.49]))
        a = oracle.smooth(x, chrom, reg, 4.0, 2.0, trim)
        b = ref.smooth(x, chrom, reg, 4.0, 2.0, trim)
        assert np.array_equal(a, b, equal_nan=True)
        fired += int(np.nansum(a != x) > 0)
    assert fired > 10  # the replacement arithmetic is exercised (the reference's own fixtures never fire it)
    for t in (-0.1,):
        for eng in (oracle, ref):
            with pytest.raises(ValueError):
                eng.smooth(np.arange(10.0), np.ones(10, np.int32), 10, 4.0, 2.0, t)
    for eng in (oracle, ref):
        with pytest.raises(OverflowError):
            eng.smooth(np.arange(10.0), np.ones(10, np.int32), 10, 4.0, 2.0, 0.0)


def test_cohort_loop_matches_reference(oracle, ref):
    from genomic_b200 import synth
    vals, off, lab, ids = synth.cohort([0], scale=0.004, outliers=True)
    for chain in (True, False):
        p = SegParams(nperm=200, chain=chain)
        a = oracle.segment_units(vals.astype(np.float64), off, lab, p)
        b = ref.segment_units(vals.astype(np.float64), off, lab, p, nthreads=1 if chain else 4)
        assert np.array_equal(a["seg_count"], b["seg_count"])
        assert np.array_equal(a["lengths"], b["lengths"]) and np.array_equal(a["means"], b["means"])


# ---- weighted CBS (cbs_oracle_weighted.c) ------------------------------------------------------------------------
CASE2_X = np.array([0.0] * 15 + [2.0] * 15 + [-1.5] * 15 + [0.0] * 15)
CASE2_W = np.array([1.0] * 15 + [0.5] * 15 + [2.0] * 15 + [1.0] * 15)


@pytest.mark.parametrize("alpha,nperm,hybrid,min_width", [(0.01, 200, False, 2), (0.05, 100, False, 3),
                                                          (0.01, 200, True, 2), (0.05, 100, True, 3)])
def test_kat_case2_segment_weighted(oracle, alpha, nperm, hybrid, min_width):
    # tests/cbs_test.cpp:309-330 (inputs literal in tests/cbs_generate.R:91-92): 15/15/15/15, means 0/2/-1.5/0
    lengths, means = oracle.segment_weighted(CASE2_X, CASE2_W, SegParams(alpha=alpha, nperm=nperm, hybrid=hybrid,
                                                                         min_width=min_width, seed=1))
    assert lengths.tolist() == [15, 15, 15, 15]
    assert np.allclose(means, [0.0, 2.0, -1.5, 0.0], atol=1e-9)


def test_segment_weighted_matches_reference(oracle, ref):
    from helpers import make_unit
    rng = np.random.default_rng(81)
    for trial in range(60):
        n = int(rng.integers(4, 1200))
        x = make_unit(rng, n, int(rng.integers(0, 5)))
        w = [rng.uniform(0.5, 2.0, n), rng.choice([0.5, 1.0, 2.0], n), np.ones(n)][trial % 3]
        p = SegParams(nperm=int(rng.choice([20, 100, 400])), alpha=float(rng.choice([0.01, 0.05])),
                      min_width=int(rng.choice([2, 3, 5])), seed=int(rng.integers(1, 100)), undo_prune=bool(trial % 8 == 7))
        if n < 2 * p.min_width:
            continue
        eng = ref.rng(p.seed)
        wl, wm = ref.segment_weighted(x, w, p, eng)
        orng = oracle.rng_mt(p.seed)
        gl, gm = oracle.segment_weighted(x, w, p, orng)
        assert np.array_equal(gl, wl), (trial, n)
        assert np.array_equal(gm, wm), (trial, n)
        assert ref.rng_equals(eng, p.seed, orng.draws), (trial, n)


def test_segment_weighted_hybrid_matches_reference(oracle, ref):
    # getmncwt / hwtmaxp / the hybrid branch of wfindcpt (CBS.cpp:593-608, 745-828, 908-921)
    from helpers import make_unit
    rng = np.random.default_rng(91)
    for trial in range(30):
        n = int(rng.integers(201, 1500))
        x = make_unit(rng, n, int(rng.integers(0, 5)))
        w = [rng.uniform(0.5, 2.0, n), rng.choice([0.5, 1.0, 2.0], n), np.ones(n)][trial % 3]
        p = SegParams(nperm=int(rng.choice([20, 100, 400])), alpha=float(rng.choice([0.01, 0.05])),
                      min_width=int(rng.choice([2, 3, 5])), seed=int(rng.integers(1, 100)), hybrid=True,
                      kmax=int(rng.choice([25, 10])), nmin=200)
        eng = ref.rng(p.seed)
        wl, wm = ref.segment_weighted(x, w, p, eng)
        orng = oracle.rng_mt(p.seed)
        gl, gm = oracle.segment_weighted(x, w, p, orng)
        assert np.array_equal(gl, wl), (trial, n)
        assert np.array_equal(gm, wm), (trial, n)
        assert ref.rng_equals(eng, p.seed, orng.draws), (trial, n)


def test_summarize_cn_golden(oracle):
    """the restatement of cngpld::summarize_cn (oracle/cngpld_oracle.c) reproduces the reference's four
    golden cases (tests/golden/cngpld, from /root/reference/tests/data via tests/cngpld_test.cpp:46-83) within the
    reference's own tolerance, and raises where the reference throws."""
    import cngpld_cases as cc
    for seg, direction, expected, positions in cc.CASES:
        s, e, v = cc.read_seg(seg)
        pos, val = oracle.summarize_cn(s, e, v, direction, cc.CUTOFF, positions)
        wpos, wval = cc.read_expected(expected)
        assert np.array_equal(pos, wpos), (seg, direction)
        assert np.allclose(val, wval, rtol=cc.RTOL, atol=0), (seg, direction)
    s, e, v = cc.read_seg("cngpld_case1_input.seg")
    with pytest.raises(ValueError):
        oracle.summarize_cn(s, e, v, 0, 0.5)           # summarize.cpp:49-51
    with pytest.raises(ValueError):
        oracle.summarize_cn(e, s, v, 1, 0.5)           # start > end, summarize.cpp:59-61
    assert len(oracle.summarize_cn(s[:0], e[:0], v[:0], 1, 0.5)[0]) == 0
