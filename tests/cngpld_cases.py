"""Test helper: the reference's golden vectors of cngpld::summarize_cn (tests/golden/cngpld, copied from the reference's
tests/data) and random segment tables for the GPU-vs-oracle comparison."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cngpld")
# (input, direction, expected, explicit positions) as in /root/reference/tests/cngpld_test.cpp:46-83
CASES = [
    ("cngpld_case1_input.seg", 1, "cngpld_case1_amp_expected.tsv", None),
    ("cngpld_case1_input.seg", -1, "cngpld_case1_del_expected.tsv", None),
    ("cngpld_case2_input.seg", 1, "cngpld_case2_amp_expected.tsv", None),
    ("cngpld_case2_input.seg", -1, "cngpld_case2_del_expected.tsv", None),
    ("cngpld_case3_input.seg", 1, "cngpld_case3_amp_expected.tsv", None),
    ("cngpld_case4_input.seg", 1, "cngpld_case4_amp_expected.tsv", [50, 100, 120, 150, 200, 220, 250]),
]
CUTOFF = 0.5
RTOL = 1e-7  # BOOST_CHECK_CLOSE(got, expected, 1e-5) is a tolerance in percent


def read_seg(name):
    rows = [l.split("\t") for l in open(os.path.join(GOLD, name)).read().splitlines()[1:] if l]
    return (np.array([int(r[2]) for r in rows], np.uint64), np.array([int(r[3]) for r in rows], np.uint64),
            np.array([float(r[5]) for r in rows], np.float32))


def read_expected(name):
    rows = [l.split() for l in open(os.path.join(GOLD, name)).read().splitlines()[1:] if l]
    return np.array([int(r[0]) for r in rows], np.uint64), np.array([float(r[1]) for r in rows], np.float64)


def random_table(rng, n_units, overlapping):
    """segment table of n_units units; overlapping=False: a partition of consecutive positions as CBS writes it"""
    off, start, end, value = [0], [], [], []
    for _ in range(n_units):
        k = int(rng.integers(0, 40))
        if overlapping:
            s = rng.integers(1, 5000, k)
            e = s + rng.integers(0, 800, k)
        else:
            cuts = np.sort(rng.choice(np.arange(1, 100000), size=2 * k, replace=False)) if k else np.array([], np.int64)
            s, e = cuts[0::2], cuts[1::2]
        start += list(s); end += list(e)
        value += list(rng.normal(0, 0.8, k))
        off.append(len(start))
    return (np.array(off, np.int64), np.array(start, np.uint64), np.array(end, np.uint64), np.array(value, np.float32))
