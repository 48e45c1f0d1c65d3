"""ctypes binding of libcbs_cuda.so (include/cbs_gpu.h).

This is the Python face of the C ABI; it mirrors the reference's call surface for the hot path
(``cbs::segment``, ``cbs::smooth``, ``cbs::tmaxo``, ``cbs::tmaxp`` and the per-chromosome loop of
``Segment::segment_raw``).  There is NO CPU fallback: if the CUDA library is missing or no
device is present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcbs_cuda.so")

OK, ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_NONFINITE, ERR_OVERFLOW = range(8)
RNG_MT19937_64, RNG_PHILOX = 0, 1
F32, F64 = 0, 1
HOST, DEVICE = 0, 1

KERNEL_NAMES = ("sched", "gen", "prep", "perm", "scan", "edgeprep", "edgeperm", "means", "smooth", "shuf0", "shuf1",
                "shuf2", "shuf3", "prefix")


class CbsGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cbs_gpu status {code}: {msg}")
        self.code = code


class CParams(C.Structure):
    _fields_ = [
        ("alpha", C.c_double), ("nperm", C.c_int32), ("hybrid", C.c_int32), ("min_width", C.c_int32),
        ("kmax", C.c_int32), ("nmin", C.c_int32), ("eta", C.c_double), ("tol", C.c_double), ("ibin", C.c_int32),
        ("undo_prune", C.c_int32), ("undo_prune_cutoff", C.c_double), ("do_smooth", C.c_int32),
        ("smooth_region", C.c_int32), ("outlier_sd_scale", C.c_double), ("smooth_sd_scale", C.c_double),
        ("trim", C.c_double), ("rng_mode", C.c_int32), ("chain", C.c_int32), ("seed", C.c_uint64),
        ("first_batch", C.c_int32), ("max_batch", C.c_int32), ("record_splits", C.c_int32), ("reserved", C.c_int32),
    ]


class CSplit(C.Structure):
    _fields_ = [
        ("unit", C.c_int32), ("lo", C.c_int32), ("hi", C.c_int32), ("ostat", C.c_double), ("iseg0", C.c_int32),
        ("iseg1", C.c_int32), ("ncpt", C.c_int32), ("icpt0", C.c_int32), ("icpt1", C.c_int32),
        ("perms_run", C.c_int32), ("nrej", C.c_int32), ("exit_code", C.c_int32), ("called", C.c_int32),
        ("e_nrej0", C.c_int32), ("e_nrej1", C.c_int32), ("e_status0", C.c_int32), ("e_status1", C.c_int32),
    ]


class CResult(C.Structure):
    _fields_ = [
        ("n_units", C.c_int32), ("n_segments", C.c_int64), ("seg_offsets", C.POINTER(C.c_int64)),
        ("lengths", C.POINTER(C.c_int32)), ("means", C.POINTER(C.c_double)),
        ("draws_consumed", C.POINTER(C.c_uint64)), ("n_splits", C.c_int64), ("splits", C.POINTER(CSplit)),
        ("rounds", C.c_int32), ("perms_run", C.c_uint64), ("perm_elements", C.c_uint64), ("kernel_launches", C.c_uint64), ("ms_h2d", C.c_double),
        ("ms_smooth", C.c_double), ("ms_segment", C.c_double), ("ms_d2h", C.c_double), ("ms_call", C.c_double),
    ]


@dataclass
class Params:
    """Options of `cna segment` (src/cna_segment.hpp:67-79) + RNG selection."""

    alpha: float = 0.01
    nperm: int = 200
    hybrid: bool = False
    min_width: int = 2
    kmax: int = 25
    nmin: int = 200
    eta: float = 0.05
    tol: float = 1e-6
    ibin: bool = False
    undo_prune: bool = False
    undo_prune_cutoff: float = 0.05
    do_smooth: bool = True
    smooth_region: int = 10
    outlier_sd_scale: float = 4.0
    smooth_sd_scale: float = 2.0
    trim: float = 0.025
    rng_mode: int = RNG_MT19937_64
    chain: bool = True
    seed: int = 1
    first_batch: int = 0
    max_batch: int = 0
    record_splits: bool = False

    def c(self) -> CParams:
        return CParams(self.alpha, self.nperm, int(self.hybrid), self.min_width, self.kmax, self.nmin, self.eta,
                       self.tol, int(self.ibin), int(self.undo_prune), self.undo_prune_cutoff, int(self.do_smooth),
                       self.smooth_region, self.outlier_sd_scale, self.smooth_sd_scale, self.trim, self.rng_mode,
                       int(self.chain), self.seed, self.first_batch, self.max_batch, int(self.record_splits), 0)


@dataclass
class BatchResult:
    seg_offsets: np.ndarray
    lengths: np.ndarray
    means: np.ndarray
    draws: np.ndarray
    splits: list = field(default_factory=list)
    rounds: int = 0
    perms_run: int = 0
    perm_elems: int = 0
    kernel_launches: int = 0
    ms: dict = field(default_factory=dict)

    @property
    def seg_count(self) -> np.ndarray:
        return np.diff(self.seg_offsets).astype(np.int32)


def load_library(path: str | None = None) -> C.CDLL:
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make -C genomic_b200/csrc` (python -c 'import "
                          "__graft_entry__ as g; g.build()'). There is no CPU fallback for this path.")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.cbs_gpu_default_params.argtypes = [C.POINTER(CParams)]
    L.cbs_gpu_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.cbs_gpu_destroy.argtypes = [vp]
    L.cbs_gpu_last_error.restype = C.c_char_p
    L.cbs_gpu_last_error.argtypes = [vp]
    L.cbs_gpu_set_stream.argtypes = [vp, vp]
    L.cbs_gpu_set_profiling.argtypes = [vp, C.c_int]
    L.cbs_gpu_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.cbs_gpu_last_arc_evals.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.cbs_gpu_measure_fp64.argtypes = [vp, C.POINTER(C.c_double)]
    L.cbs_gpu_segment_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_uint64),
                                        C.c_int32, C.POINTER(CParams), C.POINTER(C.POINTER(CResult))]
    L.cbs_gpu_result_free.argtypes = [C.POINTER(CResult)]
    L.cbs_gpu_smooth.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int64, C.c_int32, C.c_double,
                                 C.c_double, C.c_double, C.POINTER(C.c_double)]
    L.cbs_gpu_segment.argtypes = [vp, C.POINTER(C.c_double), C.c_int32, C.POINTER(CParams), C.POINTER(C.c_uint64),
                                  C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_uint64)]
    L.cbs_gpu_segment_weighted.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32,
                                           C.POINTER(CParams), C.POINTER(C.c_uint64), C.c_int32, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
    L.cbs_gpu_segment_weighted_batch.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                                 C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.c_int32,
                                                 C.POINTER(CParams), C.POINTER(C.POINTER(CResult))]
    L.cbs_gpu_tmaxo.argtypes = [vp, C.POINTER(C.c_double), C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.cbs_gpu_tmaxp.argtypes = [vp, C.POINTER(C.c_double), C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                C.POINTER(C.c_double)]
    return L


EXPORTED_SYMBOLS = (
    "cbs_gpu_default_params", "cbs_gpu_create", "cbs_gpu_destroy", "cbs_gpu_last_error", "cbs_gpu_set_stream",
    "cbs_gpu_segment_batch", "cbs_gpu_result_free", "cbs_gpu_smooth", "cbs_gpu_segment", "cbs_gpu_tmaxo",
    "cbs_gpu_tmaxp", "cbs_gpu_measure_fp64", "cbs_gpu_last_kernel_ms", "cbs_gpu_set_profiling",
    "cbs_gpu_last_arc_evals", "cbs_gpu_selftest", "cbs_gpu_segment_weighted", "cbs_gpu_segment_weighted_batch",
    "cbs_gpu_fndcpt", "cbs_gpu_wfindcpt", "cbs_gpu_tpermp", "cbs_gpu_wtmaxo", "cbs_gpu_xperm", "cbs_gpu_htmaxp",
    "cbs_gpu_tailp", "cbs_gpu_btmax", "cbs_gpu_btailp", "cbs_gpu_summarize_cn",
)


class Context:
    """One context per GPU (one process per GPU)."""

    def __init__(self, device: int = 0, lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        self.h = C.c_void_p()
        dev = (C.c_int * 1)(device)
        rc = self.lib.cbs_gpu_create(dev, 1, C.byref(self.h))
        if rc != OK:
            self.h = C.c_void_p()
            raise CbsGpuError(rc, "cbs_gpu_create failed (no CUDA device? there is no CPU fallback)")
        self.device = device

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.cbs_gpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            msg = self.lib.cbs_gpu_last_error(self.h).decode()
            if rc == ERR_INVALID:
                raise ValueError(msg)  # the reference throws std::invalid_argument
            if rc == ERR_OVERFLOW:
                raise OverflowError(msg)
            raise CbsGpuError(rc, msg)

    # ---- configuration --------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self.lib.cbs_gpu_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_profiling(self, events: bool = False, counters: bool = False, serial: bool = False):
        """events: per-launch CUDA event timing; counters: scan work counters (slow, never time with them);
        serial: all kernels on one stream, so that the per-kernel event times do not overlap."""
        self._check(self.lib.cbs_gpu_set_profiling(self.h, int(bool(events)) | (2 if counters else 0) | (4 if serial else 0)))

    def last_kernel_ms(self) -> dict:
        a = (C.c_double * 14)()
        self._check(self.lib.cbs_gpu_last_kernel_ms(self.h, a))
        return dict(zip(KERNEL_NAMES, list(a)))

    def last_arc_evals(self):
        a, s = C.c_uint64(), C.c_uint64()
        self._check(self.lib.cbs_gpu_last_arc_evals(self.h, C.byref(a), C.byref(s)))
        return a.value, s.value

    def measure_fp64(self) -> float:
        v = C.c_double()
        self._check(self.lib.cbs_gpu_measure_fp64(self.h, C.byref(v)))
        return v.value

    # ---- batched path ------------------------------------------------------------------------
    def segment_batch(self, values, unit_offsets, params: Params, unit_ids=None, device_ptr: int | None = None,
                      dtype: int | None = None) -> BatchResult:
        """values: numpy float32/float64 (host) -- or pass device_ptr + dtype for a device buffer."""
        off = np.ascontiguousarray(unit_offsets, dtype=np.int64)
        n_units = len(off) - 1
        if device_ptr is not None:
            ptr, dt, space = C.c_void_p(device_ptr), dtype, DEVICE
        else:
            if values.dtype == np.float32:
                values = np.ascontiguousarray(values)
                dt = F32
            else:
                values = np.ascontiguousarray(values, dtype=np.float64)
                dt = F64
            ptr, space = C.c_void_p(values.ctypes.data), HOST
        uid = None
        if unit_ids is not None:
            uid = np.ascontiguousarray(unit_ids, dtype=np.uint64)
        cp = params.c()
        out = C.POINTER(CResult)()
        rc = self.lib.cbs_gpu_segment_batch(self.h, ptr, dt, space, off.ctypes.data_as(C.POINTER(C.c_int64)),
                                            uid.ctypes.data_as(C.POINTER(C.c_uint64)) if uid is not None else None,
                                            n_units, C.byref(cp), C.byref(out))
        self._check(rc)
        try:
            return self._unpack(out.contents, n_units)
        finally:
            self.lib.cbs_gpu_result_free(out)

    @staticmethod
    def _unpack(r, n_units) -> BatchResult:
        ns = int(r.n_segments)
        res = BatchResult(
            seg_offsets=np.ctypeslib.as_array(r.seg_offsets, shape=(n_units + 1,)).copy(),
            lengths=np.ctypeslib.as_array(r.lengths, shape=(ns,)).copy() if ns else np.zeros(0, np.int32),
            means=np.ctypeslib.as_array(r.means, shape=(ns,)).copy() if ns else np.zeros(0),
            draws=np.ctypeslib.as_array(r.draws_consumed, shape=(n_units,)).copy() if n_units else np.zeros(0, np.uint64),
            rounds=int(r.rounds), perms_run=int(r.perms_run), perm_elems=int(r.perm_elements), kernel_launches=int(r.kernel_launches),
            ms=dict(h2d=r.ms_h2d, smooth=r.ms_smooth, segment=r.ms_segment, d2h=r.ms_d2h, call=r.ms_call),
        )
        if r.n_splits:
            res.splits = [dict((f, getattr(r.splits[i], f)) for f, _ in CSplit._fields_) for i in range(int(r.n_splits))]
        return res

    # ---- single-call surface (reference signatures) -----------------------------------------------
    def smooth(self, values, chrom, smooth_region=10, outlier_sd_scale=4.0, smooth_sd_scale=2.0, trim=0.025):
        """cbs::smooth (lib/cbs/smooth.hpp:8-13)."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        chrom = np.ascontiguousarray(chrom, dtype=np.int32)
        if len(values) != len(chrom):
            raise ValueError("values and chrom must have same length")  # smooth.cpp:125
        out = np.empty_like(values)
        self._check(self.lib.cbs_gpu_smooth(self.h, values.ctypes.data_as(C.POINTER(C.c_double)),
                                            chrom.ctypes.data_as(C.POINTER(C.c_int32)), len(values), smooth_region,
                                            outlier_sd_scale, smooth_sd_scale, trim,
                                            out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def segment(self, x, params: Params, mt_next312=None):
        """cbs::segment (lib/cbs/CBS.hpp:100-113) -> (lengths, means, draws_consumed)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        cap = max(16, len(x))
        lengths = np.zeros(cap, dtype=np.int32)
        means = np.zeros(cap, dtype=np.float64)
        nseg, draws = C.c_int32(0), C.c_uint64(0)
        cp = params.c()
        st = None
        if mt_next312 is not None:
            st = np.ascontiguousarray(mt_next312, dtype=np.uint64)
            assert len(st) == 312
        self._check(self.lib.cbs_gpu_segment(self.h, x.ctypes.data_as(C.POINTER(C.c_double)), len(x), C.byref(cp),
                                             st.ctypes.data_as(C.POINTER(C.c_uint64)) if st is not None else None,
                                             cap, lengths.ctypes.data_as(C.POINTER(C.c_int32)),
                                             means.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nseg),
                                             C.byref(draws)))
        k = nseg.value
        return lengths[:k].copy(), means[:k].copy(), draws.value

    def segment_weighted(self, x, weights, params: Params, mt_next312=None):
        """cbs::segment_weighted (lib/cbs/CBS.hpp:115-128) -> (lengths, means, draws_consumed)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if len(x) != len(w):
            raise ValueError("x and weights must have the same length")
        cap = max(16, len(x))
        lengths = np.zeros(cap, dtype=np.int32)
        means = np.zeros(cap, dtype=np.float64)
        nseg, draws = C.c_int32(0), C.c_uint64(0)
        cp = params.c()
        st = None
        if mt_next312 is not None:
            st = np.ascontiguousarray(mt_next312, dtype=np.uint64)
            assert len(st) == 312
        dp = C.POINTER(C.c_double)
        self._check(self.lib.cbs_gpu_segment_weighted(
            self.h, x.ctypes.data_as(dp), w.ctypes.data_as(dp), len(x), C.byref(cp),
            st.ctypes.data_as(C.POINTER(C.c_uint64)) if st is not None else None, cap,
            lengths.ctypes.data_as(C.POINTER(C.c_int32)), means.ctypes.data_as(dp), C.byref(nseg), C.byref(draws)))
        k = nseg.value
        return lengths[:k].copy(), means[:k].copy(), draws.value

    def segment_weighted_batch(self, values, weights, unit_offsets, params: Params, unit_ids=None) -> BatchResult:
        """cbs::segment_weighted for every unit of a flat float64 array (host memory); no smoothing."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        if len(values) != len(weights):
            raise ValueError("values and weights must have the same length")
        off = np.ascontiguousarray(unit_offsets, dtype=np.int64)
        ids = None if unit_ids is None else np.ascontiguousarray(unit_ids, dtype=np.uint64)
        cp = params.c()
        res = C.POINTER(CResult)()
        dp = C.POINTER(C.c_double)
        self._check(self.lib.cbs_gpu_segment_weighted_batch(
            self.h, values.ctypes.data_as(dp), weights.ctypes.data_as(dp), 0, off.ctypes.data_as(C.POINTER(C.c_int64)),
            ids.ctypes.data_as(C.POINTER(C.c_uint64)) if ids is not None else None, len(off) - 1, C.byref(cp),
            C.byref(res)))
        try:
            return self._unpack(res.contents, len(off) - 1)
        finally:
            self.lib.cbs_gpu_result_free(res)

    def tmaxo(self, x, tss, al0=2, ibin=False):
        """cbs::tmaxo (CBS.hpp:32) -> (statistic, start, end)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        stat, s, e = C.c_double(), C.c_int32(), C.c_int32()
        self._check(self.lib.cbs_gpu_tmaxo(self.h, x.ctypes.data_as(C.POINTER(C.c_double)), len(x), tss, al0,
                                           int(ibin), C.byref(stat), C.byref(s), C.byref(e)))
        return stat.value, s.value, e.value

    def tmaxp(self, px, tss, al0=2, ibin=False):
        """cbs::tmaxp (CBS.hpp:33) for one vector (1-D) or a stack of vectors (2-D, one per row)."""
        px = np.ascontiguousarray(px, dtype=np.float64)
        single = px.ndim == 1
        m = px.reshape(1, -1) if single else px
        out = np.zeros(m.shape[0], dtype=np.float64)
        self._check(self.lib.cbs_gpu_tmaxp(self.h, m.ctypes.data_as(C.POINTER(C.c_double)), m.shape[1], m.shape[0],
                                           tss, al0, int(ibin), out.ctypes.data_as(C.POINTER(C.c_double))))
        return float(out[0]) if single else out


    # ---- low-level call surface (lib/cbs/CBS.hpp:29-98) -------------------------------------------------------------
    @staticmethod
    def _dp(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    @staticmethod
    def _state(mt_next312):
        if mt_next312 is None:
            return None, None
        st = np.ascontiguousarray(mt_next312, dtype=np.uint64)
        assert st.shape == (312,)
        return st, st.ctypes.data_as(C.POINTER(C.c_uint64))

    @staticmethod
    def _decision(s: CSplit, draws: int) -> dict:
        return dict(ncpt=s.ncpt, icpt=(s.icpt0, s.icpt1), iseg=(s.iseg0, s.iseg1), ostat=s.ostat, perms_run=s.perms_run,
                    nrej=s.nrej, exit_code=s.exit_code, e_nrej=(s.e_nrej0, s.e_nrej1), e_status=(s.e_status0, s.e_status1),
                    draws=draws)

    def fndcpt(self, x, tss, params: Params, delta=0.0, ngrid=100, mt_next312=None):
        """cbs::fndcpt (CBS.hpp:68-80) on a centred segment: dict(ncpt, icpt, iseg, ostat, ..., draws)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        keep, st = self._state(mt_next312)
        cp, out, draws = params.c(), CSplit(), C.c_uint64(0)
        self._check(self.lib.cbs_gpu_fndcpt(self.h, self._dp(x), len(x), C.c_double(tss), C.byref(cp), C.c_double(delta), ngrid, st,
                                            C.byref(out), C.byref(draws)))
        return self._decision(out, draws.value)

    def wfindcpt(self, x, weights, tss, params: Params, ngrid=100, mt_next312=None):
        """cbs::wfindcpt (CBS.hpp:81-97); rwts and cwts are derived from weights as cbs::segment_weighted does."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        keep, st = self._state(mt_next312)
        cp, out, draws = params.c(), CSplit(), C.c_uint64(0)
        self._check(self.lib.cbs_gpu_wfindcpt(self.h, self._dp(x), self._dp(w), len(x), C.c_double(tss), C.byref(cp), ngrid, st,
                                              C.byref(out), C.byref(draws)))
        return self._decision(out, draws.value)

    def tpermp(self, n1, n2, x, params: Params, mt_next312=None):
        """cbs::tpermp (CBS.hpp:35-36) -> (p-value, draws consumed)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert len(x) >= n1 + n2
        keep, st = self._state(mt_next312)
        cp, p, draws = params.c(), C.c_double(), C.c_uint64(0)
        self._check(self.lib.cbs_gpu_tpermp(self.h, self._dp(x), n1, n2, C.byref(cp), st, C.byref(p), C.byref(draws)))
        return p.value, draws.value

    def wtmaxo(self, x, weights, tss, al0=2):
        """cbs::wtmaxo (CBS.hpp:54-58) -> (statistic, start, end)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        stat, s, e = C.c_double(), C.c_int32(), C.c_int32()
        self._check(self.lib.cbs_gpu_wtmaxo(self.h, self._dp(x), self._dp(w), len(x), C.c_double(tss), al0, C.byref(stat),
                                            C.byref(s), C.byref(e)))
        return stat.value, s.value, e.value

    def xperm(self, x, seed=1, mt_next312=None, rwts=None):
        """cbs::xperm (CBS.hpp:37), or cbs::wxperm (CBS.hpp:39-42) when rwts is given: one permutation, n draws."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        rw = None if rwts is None else np.ascontiguousarray(rwts, dtype=np.float64)
        keep, st = self._state(mt_next312)
        px = np.zeros(len(x), dtype=np.float64)
        self._check(self.lib.cbs_gpu_xperm(self.h, self._dp(x), None if rw is None else self._dp(rw), len(x), st, C.c_uint64(seed),
                                           self._dp(px)))
        return px

    def htmaxp(self, px, tss, k, al0=2, ibin=False):
        """cbs::htmaxp (CBS.hpp:34) for one vector (1-D) or a stack of vectors (2-D, one per row)."""
        px = np.ascontiguousarray(px, dtype=np.float64)
        single = px.ndim == 1
        m = px.reshape(1, -1) if single else px
        out = np.zeros(m.shape[0], dtype=np.float64)
        self._check(self.lib.cbs_gpu_htmaxp(self.h, self._dp(m), m.shape[1], m.shape[0], C.c_double(tss), k, al0, int(ibin),
                                            self._dp(out)))
        return float(out[0]) if single else out

    def tailp(self, b, delta, m, ngrid=100, tol=1e-6):
        out = C.c_double()
        self._check(self.lib.cbs_gpu_tailp(self.h, C.c_double(b), C.c_double(delta), m, ngrid, C.c_double(tol), C.byref(out)))
        return out.value

    def btmax(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = C.c_double()
        self._check(self.lib.cbs_gpu_btmax(self.h, self._dp(x), len(x), C.byref(out)))
        return out.value

    def btailp(self, b, m, ng, tol=1e-6):
        out = C.c_double()
        self._check(self.lib.cbs_gpu_btailp(self.h, C.c_double(b), m, ng, C.c_double(tol), C.byref(out)))
        return out.value

    def summarize_cn(self, seg_offsets, start, end, value, direction, cutoff, pos_offsets=None, positions=None):
        """cngpld::summarize_cn (lib/cngpld/summarize.cpp:77-100) for every unit of a segment table.
        Returns (out_offsets [n_units+1], positions uint64, values float64)."""
        so = np.ascontiguousarray(seg_offsets, np.int64)
        n_units = len(so) - 1
        st = np.ascontiguousarray(start, np.uint64)
        en = np.ascontiguousarray(end, np.uint64)
        va = np.ascontiguousarray(value, np.float32)
        if positions is not None:
            po = np.ascontiguousarray(pos_offsets, np.int64)
            ps = np.ascontiguousarray(positions, np.uint64)
            cap = int(po[-1]) if len(po) else 0
            po_p, ps_p = po.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p)
        else:
            cap = 2 * len(st)
            po_p = ps_p = None
        out_off = np.zeros(n_units + 1, np.int64)
        out_pos = np.zeros(max(cap, 1), np.uint64)
        out_val = np.zeros(max(cap, 1), np.float64)
        self._check(self.lib.cbs_gpu_summarize_cn(
            self.h, so.ctypes.data_as(C.c_void_p), n_units, st.ctypes.data_as(C.c_void_p), en.ctypes.data_as(C.c_void_p),
            va.ctypes.data_as(C.c_void_p), int(direction), C.c_double(cutoff), po_p, ps_p, out_off.ctypes.data_as(C.c_void_p),
            out_pos.ctypes.data_as(C.c_void_p), out_val.ctypes.data_as(C.c_void_p)))
        n = int(out_off[-1])
        return out_off, out_pos[:n].copy(), out_val[:n].copy()
