"""SURVEY 8 row f4 on the GPU: cngpld::summarize_cn over whole segment tables (cbs_gpu_summarize_cn) against the
reference's golden vectors and the oracle restatement."""
import numpy as np
import pytest

import cngpld_cases as cc
pytestmark = pytest.mark.gpu
RTOL = 1e-14  # CUDA exp is within 1 ulp of the host libm's; the sum and the division are the same operations


def test_summarize_cn_golden(ctx):
    for seg, direction, expected, positions in cc.CASES:
        s, e, v = cc.read_seg(seg)
        off = np.array([0, len(s)], np.int64)
        if positions is None:
            got_off, pos, val = ctx.summarize_cn(off, s, e, v, direction, cc.CUTOFF)
        else:
            got_off, pos, val = ctx.summarize_cn(off, s, e, v, direction, cc.CUTOFF, [0, len(positions)], positions)
        wpos, wval = cc.read_expected(expected)
        assert list(got_off) == [0, len(wpos)]
        assert np.array_equal(pos, wpos), (seg, direction)
        assert np.allclose(val, wval, rtol=cc.RTOL, atol=0), (seg, direction)


@pytest.mark.parametrize("overlapping", [False, True])
def test_summarize_cn_tables_match_oracle(ctx, oracle, overlapping):
    rng = np.random.default_rng(5 + overlapping)
    for trial in range(6):
        off, s, e, v = cc.random_table(rng, int(rng.integers(1, 60)), overlapping)
        for direction, cutoff in ((1, 0.5), (-1, 0.1), (1, -10.0)):
            got_off, pos, val = ctx.summarize_cn(off, s, e, v, direction, cutoff)
            for u in range(len(off) - 1):
                a, b = off[u], off[u + 1]
                wpos, wval = oracle.summarize_cn(s[a:b], e[a:b], v[a:b], direction, cutoff)
                assert np.array_equal(pos[got_off[u]:got_off[u + 1]], wpos), (trial, u)
                assert np.allclose(val[got_off[u]:got_off[u + 1]], wval, rtol=RTOL, atol=0), (trial, u)
        # explicit positions, some outside every segment
        poff, plist = [0], []
        for u in range(len(off) - 1):
            plist += list(rng.integers(0, 6000 if overlapping else 110000, int(rng.integers(0, 9))))
            poff.append(len(plist))
        got_off, pos, val = ctx.summarize_cn(off, s, e, v, 1, 0.3, poff, np.array(plist, np.uint64))
        assert list(got_off) == poff and np.array_equal(pos, np.array(plist, np.uint64))
        for u in range(len(off) - 1):
            a, b = off[u], off[u + 1]
            _, wval = oracle.summarize_cn(s[a:b], e[a:b], v[a:b], 1, 0.3, plist[poff[u]:poff[u + 1]])
            assert np.allclose(val[poff[u]:poff[u + 1]], wval, rtol=RTOL, atol=0)


def test_summarize_cn_on_a_segmentation(ctx, oracle):
    """the table CBS itself writes: segment a small cohort, turn lengths into positions, summarise amplifications"""
    from genomic_b200 import Params, synth
    vals, off, lab, ids = synth.cohort([0, 1], scale=0.01)
    r = ctx.segment_batch(vals, off, Params(nperm=200, chain=False), unit_ids=ids)
    start, end = [], []
    for u in range(len(off) - 1):
        at = 0
        for k in range(r.seg_offsets[u], r.seg_offsets[u + 1]):
            start.append(1000 * (at + 1)); end.append(1000 * (at + int(r.lengths[k]))); at += int(r.lengths[k])
    got_off, pos, val = ctx.summarize_cn(r.seg_offsets, start, end, r.means.astype(np.float32), 1, 0.05)
    assert got_off[-1] == 2 * len(start)  # a partition: every start and end is its own position
    for u in range(len(off) - 1):
        a, b = r.seg_offsets[u], r.seg_offsets[u + 1]
        wpos, wval = oracle.summarize_cn(start[a:b], end[a:b], r.means[a:b].astype(np.float32), 1, 0.05)
        assert np.array_equal(pos[got_off[u]:got_off[u + 1]], wpos)
        assert np.allclose(val[got_off[u]:got_off[u + 1]], wval, rtol=RTOL, atol=0)


def test_summarize_cn_errors(ctx):
    s, e, v = cc.read_seg("cngpld_case1_input.seg")
    off = [0, len(s)]
    with pytest.raises(ValueError):
        ctx.summarize_cn(off, s, e, v, 0, 0.5)   # direction must be 1 or -1 (summarize.cpp:49-51)
    with pytest.raises(ValueError):
        ctx.summarize_cn(off, e, s, v, 1, 0.5)   # start > end (summarize.cpp:59-61)
    got_off, pos, val = ctx.summarize_cn([0, 0, 0], s[:0], e[:0], v[:0], 1, 0.5)
    assert list(got_off) == [0, 0, 0] and len(pos) == 0
