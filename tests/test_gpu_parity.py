"""-m gpu: the CUDA path (through the C ABI, genomic_b200.Context) against the oracle on the same
seeded inputs.  Integer outputs (segment counts, lengths = breakpoints, arc locations, draws
consumed) must be identical; in MT replay mode the floating point outputs are bit-identical too
(the device sums are strictly sequential and no FMA is contracted); the north star asks 1e-9."""
import numpy as np
import pytest

from helpers import f32, make_unit, pack
from oracle.pyoracle import SegParams
import genomic_b200
from genomic_b200 import Params, RNG_MT19937_64, RNG_PHILOX

pytestmark = pytest.mark.gpu
RTOL = 1e-9  # north-star tolerance for means and t statistics


def gparams(p: SegParams, **kw) -> Params:
    return Params(alpha=p.alpha, nperm=p.nperm, hybrid=p.hybrid, min_width=p.min_width, kmax=p.kmax, nmin=p.nmin,
                  eta=p.eta, tol=p.tol, do_smooth=p.do_smooth, smooth_region=p.smooth_region,
                  outlier_sd_scale=p.outlier_sd_scale, smooth_sd_scale=p.smooth_sd_scale, trim=p.trim,
                  rng_mode=RNG_PHILOX if p.rng_kind else RNG_MT19937_64, chain=p.chain, seed=p.seed, **kw)


def check_batch(ctx, oracle, vals, off, p: SegParams, unit_ids=None, **kw):
    lab = np.ones(len(off) - 1, np.int32)
    want = oracle.segment_units(vals, off, lab, p, unit_ids=unit_ids)
    got = ctx.segment_batch(vals, off, gparams(p, **kw), unit_ids=unit_ids)
    assert np.array_equal(got.seg_count, want["seg_count"])
    assert np.array_equal(got.lengths, want["lengths"])
    assert np.allclose(got.means, want["means"], rtol=RTOL, atol=0)
    if p.rng_kind == 0:
        assert np.array_equal(got.means, want["means"])  # bit-exact in replay mode
        assert np.array_equal(got.draws, want["draws"])
    return got, want


def test_tmaxo_kat_case1(ctx):
    # tests/cbs_test.cpp:154-177 (inputs literal in tests/cbs_generate.R:90-92)
    x = np.array([0.0] * 20 + [1.5] * 20 + [0.0] * 20)
    tss = float((x * x).sum() - x.sum() ** 2 / len(x))
    stat, s, e = ctx.tmaxo(x, tss, 2, False)
    assert (s, e) == (0, 58)
    assert stat == 27000.0


def test_tmaxo_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(11)
    for trial in range(120):
        n = int(rng.integers(4, 4000))
        x = make_unit(rng, n, trial % 5)
        xc = x - x.mean()
        tss = float((xc * xc).sum())
        for al0 in (2, 3, 5):
            if n < 2 * al0:
                continue
            assert ctx.tmaxo(xc, tss, al0) == oracle.tmaxo(xc, tss, al0), (trial, n, al0)


def test_tmaxo_large(ctx, oracle):
    rng = np.random.default_rng(12)
    for n in (20000, 147726):
        x = f32(rng.normal(0, 0.2, n))
        xc = x - x.mean()
        tss = float((xc * xc).sum())
        assert ctx.tmaxo(xc, tss, 2) == oracle.tmaxo(xc, tss, 2)
        x[n // 3: n // 2] += 0.4
        xc = x - x.mean()
        tss = float((xc * xc).sum())
        assert ctx.tmaxo(xc, tss, 2) == oracle.tmaxo(xc, tss, 2)


def test_tmaxo_very_long_vector_coarse_table(ctx, oracle):
    # > 524288 markers: the scan's extrema table covers runs of 64 prefix sums per entry instead of 32
    rng = np.random.default_rng(13)
    n = 600000
    x = f32(rng.normal(0, 0.2, n))
    x[n // 5: n // 4] += 0.02
    xc = x - x.mean()
    tss = float((xc * xc).sum())
    assert ctx.tmaxo(xc, tss, 2) == oracle.tmaxo(xc, tss, 2)


def test_tmaxp_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(13)
    for n in (4, 7, 49, 50, 51, 200, 1000, 5000):
        m = np.stack([f32(rng.normal(0, 0.2, n)) for _ in range(40)])
        m -= m.mean(axis=1, keepdims=True)
        tss = float((m[0] * m[0]).sum())
        got = ctx.tmaxp(m, tss, 2)
        want = np.array([oracle.tmaxp(r, tss, 2) for r in m])
        assert np.array_equal(got, want), n


@pytest.mark.parametrize("mode", ["mt_chain", "mt_unit", "philox"])
def test_segment_batch_small_units(ctx, oracle, mode):
    rng = np.random.default_rng({"mt_chain": 1, "mt_unit": 2, "philox": 3}[mode])
    for trial in range(12):
        units = [make_unit(rng, int(rng.integers(1, 1500)), int(rng.integers(0, 5))) for _ in range(int(rng.integers(1, 7)))]
        if trial % 4 == 0:
            units.insert(1, np.zeros(0))  # empty chromosome is skipped (cna_segment.hpp:138)
        vals, off = pack(units)
        p = SegParams(nperm=int(rng.choice([50, 200, 1000])), alpha=float(rng.choice([0.01, 0.05])),
                      min_width=int(rng.choice([2, 3])), do_smooth=False, rng_kind=1 if mode == "philox" else 0,
                      chain=(mode == "mt_chain"), seed=int(rng.integers(1, 100)))
        check_batch(ctx, oracle, vals, off, p, first_batch=int(rng.choice([16, 64, 256])), max_batch=int(rng.choice([256, 2048])))


def test_split_log_matches(ctx, oracle):
    rng = np.random.default_rng(21)
    units = [make_unit(rng, 900, k) for k in (0, 1, 1, 3)]
    vals, off = pack(units)
    p = SegParams(nperm=500, alpha=0.01, do_smooth=False, chain=True, seed=1)
    want = oracle.segment_units(vals, off, np.ones(4, np.int32), p, want_log=True)
    got = ctx.segment_batch(vals, off, gparams(p, record_splits=True))
    assert len(got.splits) == len(want["log"])
    for g, (u, w) in zip(got.splits, want["log"]):
        assert (g["unit"], g["lo"], g["hi"], g["called"], g["ncpt"]) == (u, w.lo, w.hi, w.called, w.ncpt)
        assert g["ostat"] == w.ostat  # t statistic, bit-exact (tolerance allowed: 1e-9)
        if w.called:
            assert (g["iseg0"], g["iseg1"], g["perms_run"], g["exit_code"]) == (w.iseg0, w.iseg1, w.perms_run, w.exit_code)
            if w.ncpt >= 1:
                assert g["icpt0"] == w.icpt0
            if w.ncpt == 2:
                assert g["icpt1"] == w.icpt1
            check_edge_pvalues(g, w, p.nperm)


def check_edge_pvalues(g, w, nperm):
    """tpermp p-values of both boundaries (CBS.cpp:877,883): the device reports the rejection counts / shortcut status"""
    for side, want_p in ((0, w.edge_p0), (1, w.edge_p1)):
        if want_p < 0:
            continue  # tpermp did not run (no split, or a single change point at the segment end)
        st, nrej = g[f"e_status{side}"], g[f"e_nrej{side}"]
        got_p = 1.0 if st == 1 else 0.0 if st == 2 else nrej / nperm
        assert got_p == want_p, (g["unit"], g["lo"], g["hi"], side, st, nrej, want_p)


def check_split_log(got, want, nperm):
    assert len(got.splits) == len(want["log"])
    compared = 0
    by_key = {(g["unit"], g["lo"], g["hi"]): g for g in got.splits}  # decisions are logged in order of completion
    assert len(by_key) == len(got.splits)
    for u, w in want["log"]:
        g = by_key[(u, w.lo, w.hi)]
        assert (g["called"], g["ncpt"]) == (w.called, w.ncpt)
        if w.called:
            assert (g["perms_run"], g["exit_code"]) == (w.perms_run, w.exit_code)
            check_edge_pvalues(g, w, nperm)
            compared += int(0.0 < w.edge_p0 < 1.0) + int(0.0 < w.edge_p1 < 1.0)
    return compared


def test_edge_pvalues_chain_mode(ctx, oracle):
    """One serial stream (chain=1, what `cna segment` does), weak shifts in the middle of units of 4000-7000 markers and a
    generous alpha, so that tpermp runs on many boundaries, with up to 1400 draws per permutation (nperm x m1 ~ 3e6 per test) and p-values strictly between 0 and 1.  The edge permutations must read THIS round's draws, i.e. run after the
    generator: every tpermp rejection count is compared with the oracle (lengths alone can hide a stale read)."""
    rng = np.random.default_rng(131)
    compared = 0
    for trial in range(4):
        units = []
        for k in range(2):
            n = int(rng.integers(4000, 7000))
            x = rng.normal(0, 0.2, n)
            a = int(rng.integers(n // 5, n // 3)); b = int(rng.integers(n // 2, 4 * n // 5))
            x[a:b] += float(rng.choice([0.015, 0.02]))
            units.append(f32(x))
        vals, off = pack(units)
        p = SegParams(nperm=2000, alpha=0.6, do_smooth=False, rng_kind=0, chain=True, seed=1 + trial)
        want = oracle.segment_units(vals, off, np.ones(len(off) - 1, np.int32), p, want_log=True)
        got = ctx.segment_batch(vals, off, gparams(p, record_splits=True))
        assert np.array_equal(got.lengths, want["lengths"]) and np.array_equal(got.draws, want["draws"])
        compared += check_split_log(got, want, p.nperm)
    assert compared >= 20  # permutation-based edge tests with 0 < p < 1 were compared


def test_edge_tests_large_m1(ctx, oracle):
    # weak shifts in the middle: tpermp runs with large m1 (general edge kernel), CBS.cpp:495-536
    rng = np.random.default_rng(31)
    for trial in range(8):
        n = int(rng.integers(300, 1500))
        x = rng.normal(0, 0.2, n)
        a = int(rng.integers(n // 5, n // 2))
        b = int(rng.integers(a + 70, n - 70))
        x[a:b] += float(rng.choice([0.08, 0.1, 0.12, 0.15]))
        vals, off = pack([f32(x)])
        for mode in range(3):
            p = SegParams(nperm=300, alpha=0.05, do_smooth=False, rng_kind=1 if mode == 2 else 0, chain=(mode == 0), seed=3)
            check_batch(ctx, oracle, vals, off, p, first_batch=32, max_batch=128)
            want = oracle.segment_units(vals, off, np.ones(1, np.int32), p, want_log=True)
            got = ctx.segment_batch(vals, off, gparams(p, record_splits=True, first_batch=32, max_batch=128))
            check_split_log(got, want, p.nperm)


def test_smooth_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(41)
    for trial in range(40):
        n = int(rng.integers(0, 6000))
        x = rng.normal(0, 0.2, n)
        for i in (rng.integers(0, max(n, 1), max(1, n // 50)) if n else []):
            x[i] += rng.choice([-1, 1]) * (3 + abs(rng.normal()))
        if trial % 3 == 0 and n:
            for i in rng.integers(0, n, 3):
                x[i] = rng.choice([np.nan, np.inf, -np.inf])
        chrom = np.sort(rng.integers(1, 4, n)).astype(np.int32) if trial % 2 else np.ones(n, np.int32)
        reg = int(rng.choice([0, 1, 2, 10]))
        trim = float(rng.choice([0.025, 0.01, 0.1, 0.49]))
        got = ctx.smooth(x, chrom, reg, 4.0, 2.0, trim)
        want = oracle.smooth(x, chrom, reg, 4.0, 2.0, trim)
        assert np.array_equal(got, want, equal_nan=True), (trial, n, reg, trim)
        if n > 200 and trial % 3:
            assert (got != x).sum() > 0  # the replacement branch fires on injected outliers


def test_smooth_errors(ctx):
    x = np.arange(10.0)
    lab = np.ones(10, np.int32)
    with pytest.raises(ValueError):
        ctx.smooth(x, lab, -1)  # smooth.cpp:126
    with pytest.raises(ValueError):
        ctx.smooth(x, lab[:5])  # smooth.cpp:125
    with pytest.raises(ValueError):
        ctx.smooth(x, lab, 10, 4.0, 2.0, -0.1)  # smooth.cpp:16-18
    ctx.smooth(x, lab, 10, 4.0, 2.0, 0.7)  # reference: n_keep <= 0 returns before validating trim


def test_smooth_plus_segment_cohort(ctx, oracle):
    # config 3 shape at 1/100 scale: smoothing + CBS on units with injected outliers
    from genomic_b200 import synth
    vals, off, lab, ids = synth.cohort([0, 1], scale=0.01, outliers=True)
    for mode in range(3):
        p = SegParams(nperm=1000, alpha=0.01, do_smooth=True, rng_kind=1 if mode == 2 else 0, chain=(mode == 0), seed=1)
        want = oracle.segment_units(vals.astype(np.float64), off, lab, p, unit_ids=ids)
        got = ctx.segment_batch(vals, off, gparams(p), unit_ids=ids)  # float32 in, widened on device
        assert np.array_equal(got.seg_count, want["seg_count"])
        assert np.array_equal(got.lengths, want["lengths"])
        assert np.array_equal(got.means, want["means"])


def test_golden_cli_fixture(ctx):
    # tests/cna_test.cpp:258-270: `cna segment` on segment_cli_case1_input.cn, byte-identical .seg
    import os
    from cnio import cohort_from_cn, seg_text
    here = os.path.dirname(os.path.abspath(__file__))
    names, positions, _, values, off, lab, units = cohort_from_cn(os.path.join(here, "golden", "segment_cli_case1_input.cn"))
    got = ctx.segment_batch(values.astype(np.float32), off, Params())  # CLI defaults, MT seed 1, chain
    txt = seg_text(names, positions, units, got.seg_count, got.lengths, got.means)
    assert txt == open(os.path.join(here, "golden", "segment_cli_case1_expected.seg")).read()


def test_nonfinite_rejected(ctx):
    x = np.array([0.1, np.nan, -0.2, 0.3, 0.0, 0.5])
    with pytest.raises(genomic_b200.CbsGpuError):
        ctx.segment_batch(x, np.array([0, 6]), Params(do_smooth=False))


def test_segment_single_call_with_engine_state(ctx, oracle):
    # cbs::segment with a caller-owned engine that has already been used (rng is in/out)
    import ctypes as C
    rng = np.random.default_rng(51)
    x = make_unit(rng, 700, 1)
    eng = oracle.rng_mt(9)
    oracle.lib.orc_rng_discard(C.byref(eng), 1234)
    # next 312 raw words of that engine
    probe = oracle.rng_mt(9)
    oracle.lib.orc_rng_discard(C.byref(probe), 1234)
    outs = np.array([oracle.lib.orc_rng_u64(C.byref(probe)) for _ in range(312)], dtype=np.uint64)

    def untemper(y):
        y = int(y)
        M = (1 << 64) - 1
        y ^= y >> 43
        y ^= (y << 37) & 0xFFF7EEE000000000 & M
        z = y
        for _ in range(4):
            z = y ^ ((z << 17) & 0x71D67FFFEDA60000 & M)
        y = z
        z = y
        for _ in range(3):
            z = y ^ ((z >> 29) & 0x5555555555555555)
        return z
    nxt = np.array([untemper(v) for v in outs], dtype=np.uint64)
    p = SegParams(nperm=500, alpha=0.01, do_smooth=False, seed=9)
    wl, wm = oracle.segment(x, p, eng)
    gl, gm, draws = ctx.segment(x, gparams(p), mt_next312=nxt)
    assert np.array_equal(gl, wl) and np.array_equal(gm, wm)
    assert draws == eng.draws - 1234


@pytest.mark.parametrize("mode", ["mt_unit", "philox"])
def test_segment_long_units(ctx, oracle, mode):
    # one unit per shuffle size class, including > 65535 markers (index array in global memory)
    rng = np.random.default_rng(61)
    units = [f32(rng.normal(0, 0.2, n)) for n in (5000, 20000, 40000, 70000)]
    units[1][9000:] += 0.05
    vals, off = pack(units)
    p = SegParams(nperm=40, alpha=0.05, do_smooth=False, rng_kind=1 if mode == "philox" else 0, chain=False, seed=4)
    check_batch(ctx, oracle, vals, off, p, first_batch=24, max_batch=64)


def test_small_stream_window(oracle, monkeypatch):
    """MT replay with one engine per unit reads ONE shared raw stream through a window (ring buffer).  With the ring
    forced down to its minimum (CBS_GPU_STREAM_MB=1) the units' cursors run many windows apart: fast chains wait for
    slow ones, the ring wraps, and the result is still the oracle's, draws included.  (Round 1 materialised the whole
    stream and failed with a capacity error when a unit outran the buffer.)"""
    monkeypatch.setenv("CBS_GPU_STREAM_MB", "1")
    c = genomic_b200.Context(0)
    try:
        rng = np.random.default_rng(63)
        units = [f32(rng.normal(0, 0.2, n)) for n in (3000, 9000, 700, 20000, 5000)]
        units[1][4000:] += 0.03
        units[3][:9000] -= 0.02
        vals, off = pack(units)
        p = SegParams(nperm=2000, alpha=0.01, do_smooth=False, rng_kind=0, chain=False, seed=5)
        got, want = check_batch(c, oracle, vals, off, p)
        assert int(want["draws"].max()) > 8 * (1 << 17)  # far more draws than the 2^17..2^18-word ring holds
    finally:
        c.close()


def test_sub_batches_of_a_large_call(oracle, monkeypatch):
    """Calls beyond CBS_GPU_CHUNK_MARKERS markers are segmented in consecutive sub-batches (cbs_gpu.cu
    cbs_gpu_segment_batch) and merged: with the limit forced down to a few thousand markers the result -- segments, means,
    draws, split log, smoothing included -- is still the oracle's for Philox keys (global unit ids) and for MT replay with one
    engine per unit; MT replay with ONE chained engine is never split."""
    monkeypatch.setenv("CBS_GPU_CHUNK_MARKERS", "4000")
    c = genomic_b200.Context(0)
    try:
        rng = np.random.default_rng(77)
        units = [f32(rng.normal(0, 0.2, n)) for n in (1500, 2500, 300, 0, 5200, 900, 2100, 40, 3300)]
        units[1][1000:] += 0.5
        units[4][2000:2600] -= 0.6
        units[8][::500] += 3.0  # outliers for the smoother
        vals, off = pack(units)
        for kind, chain in ((1, False), (0, False), (0, True)):
            p = SegParams(nperm=400, alpha=0.01, do_smooth=True, rng_kind=kind, chain=chain, seed=11)
            got, want = check_batch(c, oracle, vals, off, p)
            monkeypatch.delenv("CBS_GPU_CHUNK_MARKERS")
            whole, _ = check_batch(c, oracle, vals, off, p)  # one call for all units
            monkeypatch.setenv("CBS_GPU_CHUNK_MARKERS", "4000")
            assert (got.rounds > whole.rounds) == (not chain)  # the rounds of the sub-batches add up
    finally:
        c.close()


def test_stress_50k_segments_replay(ctx, oracle):
    """BASELINE configs[4] in small: units of exactly 50,000 markers, deterministic MT replay.  (i) pure null: the
    permutation loop stops at the early exit; (ii) a shift at 25,000 small enough (t ~ 5.6 < 7) that fndcpt does not
    take the big-t shortcut, so all nperm permutations and both edge tests run.  nperm is what the CPU oracle can do
    in seconds (the full 100k-permutation case is a bench-only property run); results must be bit-identical."""
    from genomic_b200 import synth
    units = [synth.null_unit(20260105, 50000).astype(np.float64),
             synth.null_unit(20260106, 50000, shift_at=25000, shift=0.01).astype(np.float64)]
    vals, off = pack(units)
    p = SegParams(nperm=300, alpha=0.01, do_smooth=False, rng_kind=0, chain=False, seed=1)
    got, want = check_batch(ctx, oracle, vals, off, p)
    assert got.perms_run >= 300  # unit (ii) ran the whole permutation loop at least once


@pytest.mark.parametrize("mode", ["mt_chain", "mt_unit", "philox"])
def test_hybrid_matches_oracle(ctx, oracle, mode):
    """hybrid p-values (htmaxp + tailp, CBS.cpp:387-485, :324-339; tests/cbs_test.cpp:264-285 style)"""
    rng = np.random.default_rng({"mt_chain": 71, "mt_unit": 72, "philox": 73}[mode])
    for trial in range(8):
        units = [make_unit(rng, int(rng.integers(150, 2500)), int(rng.integers(0, 4))) for _ in range(int(rng.integers(1, 5)))]
        vals, off = pack(units)
        p = SegParams(nperm=int(rng.choice([100, 200, 1000])), alpha=float(rng.choice([0.01, 0.05])), hybrid=True,
                      min_width=int(rng.choice([2, 3])), kmax=int(rng.choice([25, 10])), do_smooth=False,
                      rng_kind=1 if mode == "philox" else 0, chain=(mode == "mt_chain"), seed=int(rng.integers(1, 100)))
        check_batch(ctx, oracle, vals, off, p)


def test_hybrid_long_unit(ctx, oracle):
    # 60,000 markers: the analytic tail probability (tailp/nu, CBS.cpp:18-41, :324-339) sums ~1e5 series terms per
    # quadrature point here; on the device a warp evaluates them 32 at a time and subtracts them in the reference's order
    from genomic_b200 import synth
    units = [synth.null_unit(20260107, 60000, shift_at=20000, shift=0.012).astype(np.float64)]
    vals, off = pack(units)
    p = SegParams(nperm=200, alpha=0.01, hybrid=True, do_smooth=False, rng_kind=0, chain=False, seed=1)
    check_batch(ctx, oracle, vals, off, p)


def test_hybrid_kat_case1(ctx):
    # tests/cbs_test.cpp:287-307 hybrid rows: lengths 20/20/20, means 0/1.5/0
    x = np.array([0.0] * 20 + [1.5] * 20 + [0.0] * 20)
    for alpha, nperm, mw in ((0.01, 200, 2), (0.05, 100, 3)):
        lengths, means, _ = ctx.segment(x, Params(alpha=alpha, nperm=nperm, hybrid=True, min_width=mw, do_smooth=False, seed=1))
        assert lengths.tolist() == [20, 20, 20]
        assert np.allclose(means, [0.0, 1.5, 0.0], atol=1e-9)


def test_undo_prune_matches_oracle(ctx, oracle):
    # undo.splits="prune" (CBS.cpp:266-320): weak extra steps get merged back
    rng = np.random.default_rng(81)
    hits = 0
    for trial in range(10):
        n = int(rng.integers(400, 1500))
        x = rng.normal(0, 0.2, n)
        for _ in range(int(rng.integers(2, 5))):
            a = int(rng.integers(0, n - 40)); b = int(rng.integers(a + 20, n))
            x[a:b] += float(rng.choice([0.12, 0.2, 0.5]))
        vals, off = pack([f32(x)])
        for cutoff in (0.05, 0.5):
            p = SegParams(nperm=200, alpha=0.05, do_smooth=False, undo_prune=True, undo_prune_cutoff=cutoff, seed=2)
            base = oracle.segment_units(vals, off, np.ones(1, np.int32), SegParams(nperm=200, alpha=0.05, do_smooth=False, seed=2))
            want = oracle.segment_units(vals, off, np.ones(1, np.int32), p)
            got = ctx.segment_batch(vals, off, gparams(p, undo_prune=True, undo_prune_cutoff=cutoff))
            assert np.array_equal(got.lengths, want["lengths"]) and np.array_equal(got.means, want["means"])
            hits += int(len(want["lengths"]) != len(base["lengths"]))
    assert hits > 0  # pruning actually changed something


def test_full_size_sample_properties(ctx, oracle):
    """BASELINE configs[1] at full size (1.8 M markers, 23 chromosomes, nperm 10 000, smoothing on): size-independent
    properties -- lengths partition every unit, the means are the sequential sums of the smoothed values, the call is
    deterministic, and cutting the cohort into two calls (what sharding by sample/chromosome does) changes nothing; the
    smallest chromosome is checked against the single-threaded C oracle.  The comparison of ALL units of this sample with
    the compiled reference (16 host threads, ~100 s) is tests/test_gpu_fullsize.py::test_config2_whole_sample_vs_reference."""
    from genomic_b200 import synth
    vals, off, lab, ids = synth.cohort([0], scale=1.0)
    gp = Params(nperm=10000, alpha=0.01, rng_mode=RNG_MT19937_64, chain=False, seed=1)
    a = ctx.segment_batch(vals, off, gp, unit_ids=ids)
    b = ctx.segment_batch(vals, off, gp, unit_ids=ids)
    assert np.array_equal(a.lengths, b.lengths) and np.array_equal(a.means, b.means) and np.array_equal(a.draws, b.draws)
    n_units = len(off) - 1
    x64 = vals.astype(np.float64)
    for u in range(n_units):
        s0, s1 = int(a.seg_offsets[u]), int(a.seg_offsets[u + 1])
        lens = a.lengths[s0:s1]
        assert lens.min() >= 1 and int(lens.sum()) == int(off[u + 1] - off[u])
        sm = ctx.smooth(x64[off[u]:off[u + 1]], np.full(int(off[u + 1] - off[u]), int(lab[u]), np.int32))
        pos = 0
        for k, ln in enumerate(lens):
            want = np.cumsum(sm[pos:pos + ln])[-1] / float(ln)  # np.cumsum adds strictly in order, like CBS.cpp:1019
            assert a.means[s0 + k] == want, (u, k)
            pos += int(ln)
    # two calls over disjoint unit ranges == one call (unit ids keep the Philox keys / per-unit MT streams the same)
    cutu = 12
    c1 = ctx.segment_batch(vals[:off[cutu]], off[:cutu + 1], gp, unit_ids=ids[:cutu])
    c2 = ctx.segment_batch(vals[off[cutu]:], off[cutu:] - off[cutu], gp, unit_ids=ids[cutu:])
    assert np.array_equal(np.concatenate([c1.lengths, c2.lengths]), a.lengths)
    assert np.array_equal(np.concatenate([c1.means, c2.means]), a.means)
    assert np.array_equal(np.concatenate([c1.draws, c2.draws]), a.draws)
    # smallest chromosome against the oracle (same smoothing + CBS, MT replay)
    u = int(np.argmin(np.diff(off)))
    p = SegParams(nperm=10000, alpha=0.01, do_smooth=True, rng_kind=0, chain=False, seed=1)
    want = oracle.segment_units(x64[off[u]:off[u + 1]], np.array([0, off[u + 1] - off[u]], np.int64), lab[u:u + 1], p,
                                unit_ids=ids[u:u + 1])
    s0, s1 = int(a.seg_offsets[u]), int(a.seg_offsets[u + 1])
    assert np.array_equal(a.lengths[s0:s1], want["lengths"]) and np.array_equal(a.means[s0:s1], want["means"])
    assert a.draws[u] == want["draws"][0]
