"""CPU (-m "not gpu"): host-side logic without a device.
 * libcbs_cuda.so loads and exports every symbol include/cbs_gpu.h declares; creating a context
   without a GPU fails loudly (no CPU fallback);
 * the worklist scheduler + thread-per-permutation code, compiled for the host (tests/emul), reproduce
   the oracle in all three RNG modes, under tiny batches / tiny arenas (deferral paths);
 * the MT19937-64 jump-ahead table passes its self test against sequential generation;
 * the sharded (world_size 2, gloo) gather of segment tables equals the single-process table."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import make_unit, pack
from oracle.pyoracle import SegParams, _dp, _ip, c_i64_p, c_u64_p

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_library_exports_header_symbols():
    import genomic_b200
    lib = genomic_b200.load_library()
    header = open(os.path.join(ROOT, "include", "cbs_gpu.h")).read()
    declared = set(re.findall(r"\b(cbs_gpu_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cbs_gpu.h but not exported"
    assert declared == set(genomic_b200.EXPORTED_SYMBOLS)


def test_no_cpu_fallback():
    import torch
    import genomic_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(genomic_b200.CbsGpuError):
        genomic_b200.Context(0)


def test_mt_jump_selftest():
    import genomic_b200
    lib = genomic_b200.load_library()
    assert lib.cbs_gpu_selftest() == 0


@pytest.fixture(scope="module")
def emul():
    d = os.path.join(HERE, "emul")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", os.path.join(d, "libemul.so"),
                    os.path.join(d, "emul.cpp"), "-L" + os.path.join(ROOT, "oracle"), "-l:liboracle.so",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    L = C.CDLL(os.path.join(d, "libemul.so"))
    L.emul_segment_units.restype = C.c_int64

    def run(values, off, p, first_batch=64, max_batch=512, arena_cap=1 << 22, draws_cap=1 << 22, max_live=64):
        values = np.ascontiguousarray(values, dtype=np.float64)
        off = np.ascontiguousarray(off, dtype=np.int64)
        nu = len(off) - 1
        cap = len(values) + nu + 16
        sc = np.zeros(nu, np.int32); ln = np.zeros(cap, np.int32); mn = np.zeros(cap); dr = np.zeros(nu, np.uint64)
        nr = C.c_int(0)
        tot = L.emul_segment_units(_dp(values), off.ctypes.data_as(c_i64_p), None, nu, C.c_double(p.alpha), p.nperm,
                                   p.min_width, p.rng_kind, int(p.chain), C.c_uint64(p.seed), first_batch, max_batch,
                                   C.c_longlong(arena_cap), C.c_longlong(draws_cap), max_live, C.c_int64(cap), _ip(sc), _ip(ln),
                                   _dp(mn), dr.ctypes.data_as(c_u64_p), C.byref(nr), 0, None, None)
        assert tot >= 0, tot
        return dict(seg_count=sc, lengths=ln[:tot].copy(), means=mn[:tot].copy(), draws=dr)
    return run


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_scheduler_emulation_matches_oracle(oracle, emul, mode):
    rng = np.random.default_rng(5 + mode)
    for trial in range(25):
        units = [make_unit(rng, int(rng.integers(1, 1200)), int(rng.integers(0, 5))) for _ in range(int(rng.integers(1, 6)))]
        if trial % 7 == 0:
            units.insert(1, np.zeros(0))
        vals, off = pack(units)
        p = SegParams(nperm=int(rng.choice([50, 200, 1000])), alpha=float(rng.choice([0.01, 0.05])),
                      min_width=int(rng.choice([2, 3])), do_smooth=False, rng_kind=1 if mode == 2 else 0,
                      chain=(mode == 0), seed=int(rng.integers(1, 100)))
        want = oracle.segment_units(vals, off, np.ones(len(off) - 1, np.int32), p)
        fb = int(rng.choice([3, 16, 64, 256]))
        got = emul(vals, off, p, first_batch=fb, max_batch=max(fb, int(rng.choice([8, 64, 2048]))),
                   arena_cap=int(rng.choice([1 << 16, 1 << 22])), draws_cap=int(rng.choice([1 << 16, 1 << 22])),
                   max_live=int(rng.choice([1, 2, 64])))
        assert np.array_equal(want["seg_count"], got["seg_count"])
        assert np.array_equal(want["lengths"], got["lengths"]) and np.array_equal(want["means"], got["means"])
        if mode < 2:
            assert np.array_equal(want["draws"], got["draws"])


def test_scheduler_emulation_small_stream_window(oracle, emul, monkeypatch):
    """MT replay, one engine per unit: the shared stream is a window (ring) on the engine's output.  With a ring of
    16384 words the fast chains must wait for the slow ones and the ring wraps many times; results and draw counts
    stay those of the oracle (no input may exhaust the stream buffer)."""
    monkeypatch.setenv("EMUL_STREAM_RING", "16384")
    rng = np.random.default_rng(77)
    for trial in range(10):
        units = [make_unit(rng, int(rng.integers(200, 1500)), int(rng.integers(0, 5))) for _ in range(int(rng.integers(2, 7)))]
        units.insert(int(rng.integers(0, len(units))), make_unit(rng, 1300, 0))  # a null unit: thousands of draws per permutation batch
        vals, off = pack(units)
        p = SegParams(nperm=int(rng.choice([200, 1000])), alpha=0.01, do_smooth=False, rng_kind=0, chain=False,
                      seed=int(rng.integers(1, 100)))
        want = oracle.segment_units(vals, off, np.ones(len(off) - 1, np.int32), p)
        assert int(want["draws"].max()) > 16384  # the stream really outruns the ring
        got = emul(vals, off, p, first_batch=16, max_batch=64, max_live=64)  # all chains start in round 0, as in the product
        assert np.array_equal(want["seg_count"], got["seg_count"])
        assert np.array_equal(want["lengths"], got["lengths"]) and np.array_equal(want["means"], got["means"])
        assert np.array_equal(want["draws"], got["draws"])


def test_scheduler_emulation_edge_tests(oracle, emul):
    rng = np.random.default_rng(9)
    for trial in range(12):
        n = int(rng.integers(300, 1500))
        x = rng.normal(0, 0.2, n)
        a = int(rng.integers(n // 5, n // 2)); b = int(rng.integers(a + 70, n - 70))
        x[a:b] += float(rng.choice([0.08, 0.1, 0.12, 0.15]))
        x = x.astype(np.float32).astype(np.float64)
        off = np.array([0, n])
        for mode in range(3):
            p = SegParams(nperm=300, alpha=0.05, do_smooth=False, rng_kind=1 if mode == 2 else 0, chain=(mode == 0), seed=3)
            want = oracle.segment_units(x, off, np.ones(1, np.int32), p)
            got = emul(x, off, p, first_batch=32, max_batch=128, arena_cap=1 << 20, draws_cap=1 << 20)
            assert np.array_equal(want["lengths"], got["lengths"]) and np.array_equal(want["means"], got["means"])


def test_undo_prune_host_matches_reference(ref):
    """genomic_b200/csrc/prune.h (depth-first search over tabulated group terms) against the reference's prune_segments
    (CBS.cpp:266-320) reached through cbs::segment(..., undo_prune=true): same lengths for cut-offs that merge nothing, some
    and all change points, on noisy data with true and spurious change points."""
    d = os.path.join(HERE, "cpp")
    so = os.path.join(d, "libprune_host.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, os.path.join(d, "prune_host.cpp")], check=True)
    L = C.CDLL(so)
    rng = np.random.default_rng(404)
    merged_some = 0
    for trial in range(30):
        n = int(rng.integers(60, 900))
        x = rng.normal(0, 0.2, n)
        for _ in range(int(rng.integers(1, 7))):
            a = int(rng.integers(0, n - 5)); b = int(rng.integers(a + 3, n + 1))
            x[a:b] += float(rng.choice([-0.6, -0.25, 0.2, 0.35, 0.8]))
        x = x.astype(np.float32).astype(np.float64)
        for cutoff in (0.0, 0.05, 0.5, 5.0):
            base = SegParams(nperm=200, alpha=0.05, do_smooth=False, seed=int(trial + 1))
            plain = ref.segment(x, base)
            want = ref.segment(x, SegParams(nperm=200, alpha=0.05, do_smooth=False, seed=int(trial + 1), undo_prune=True,
                                            undo_prune_cutoff=cutoff))
            lseg = np.ascontiguousarray(plain[0], dtype=np.int32)
            out = np.zeros(len(lseg) + 1, np.int32)
            k = L.prune_host(_dp(x), n, _ip(lseg), len(lseg), C.c_double(cutoff), _ip(out))
            assert np.array_equal(out[:k], want[0]), (trial, cutoff, lseg, out[:k], want[0])
            merged_some += int(1 < k < len(lseg))
    assert merged_some > 0


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from genomic_b200 import shard, synth
from oracle.pyoracle import Oracle, SegParams
rank, world = int(sys.argv[1]), 2
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[2]
dist.init_process_group("gloo", rank=rank, world_size=world)
mine = shard.samples_of_rank(3, rank, world)
vals, off, lab, ids = synth.cohort(mine, scale=0.002)
p = SegParams(nperm=100, rng_kind=1, seed=11)
r = Oracle().segment_units(vals.astype(np.float64), off, lab, p, unit_ids=ids)
tab = shard.gather_tables(shard.pack_table(r["seg_count"], r["lengths"], r["means"], ids), dist, capacity=int(sys.argv[4]))
if rank == 0:
    np.save(sys.argv[3], tab)
dist.barrier(); dist.destroy_process_group()
'''


def test_sharded_gather_gloo_world2(tmp_path, oracle):
    """the N>1 path without GPUs: two gloo ranks segment their samples (oracle as the stand-in engine) and
    gather; the gathered table must equal the single-process table (global unit ids => identical philox keys)"""
    from genomic_b200 import shard, synth
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    assert shard.samples_of_rank(3, 0, 2) == [0, 1] and shard.samples_of_rank(3, 1, 2) == [2]
    vals, off, lab, ids = synth.cohort([0, 1, 2], scale=0.002)
    r = oracle.segment_units(vals.astype(np.float64), off, lab, SegParams(nperm=100, rng_kind=1, seed=11), unit_ids=ids)
    want = shard.pack_table(r["seg_count"], r["lengths"], r["means"], ids)
    for k, capacity in enumerate((4096, 3)):  # one collective; with 3 rows per rank the gather has to be repeated with the real size
        out = tmp_path / f"tab{k}.npy"
        port = str(29500 + (os.getpid() + k) % 2000)
        procs = [subprocess.Popen([sys.executable, str(script), str(r2), port, str(out), str(capacity)]) for r2 in range(2)]
        for pr in procs:
            assert pr.wait(timeout=300) == 0
        assert np.array_equal(np.load(out), want)


def test_cn_reader_parallel_identical(tmp_path):
    """genomic_b200/host/cn_reader.hpp: the multi-threaded .cn reader returns exactly what the sequential restatement of
    RawSampleSet<float>::_read returns (unknown chromosomes, unsorted positions, nan fields, unterminated last line)"""
    import json
    import subprocess
    exe = tmp_path / "reader_test"
    subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", os.path.join(ROOT, "tests", "cpp", "reader_test.cpp"), "-o", str(exe)],
                   check=True)
    for markers, samples, threads in ((37, 3, 4), (1000, 7, 5)):
        r = subprocess.run([str(exe), str(tmp_path / "t.cn"), str(markers), str(samples), str(threads)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert json.loads(r.stdout)["identical"] is True


def test_cn_reader_reference_rules(tmp_path):
    """tests/cpp/reader_rules_test.cpp: adversarial .cn inputs (rows the reference drops, fields it skips and the column
    shift that follows, nan/inf, rows sharing a position in std::sort's order, unterminated last line, empty files)
    against expectations derived from lib/RawSampleSet.hpp:217-285,332-386 and lib/parse.hpp:20-26"""
    import subprocess
    exe = tmp_path / "reader_rules_test"
    subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", os.path.join(ROOT, "tests", "cpp", "reader_rules_test.cpp"), "-o", str(exe)],
                   check=True)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
