// binary.cuh -- the binary-data variants on the lib/cbs call surface: cbs::tmaxo / cbs::tmaxp with ibin = true
// (CBS.cpp:68-227, the ibin branches), cbs::btmax (:363-376) and cbs::btailp (:341-361).
//
// `cna segment` always passes ibin = false (src/cna_segment.hpp:141), so these are OFF the hot path; they exist so that the
// reference's own tests of this surface (tests/cbs_test.cpp:154-177) pass under the namespace swap.  With the continuity
// correction the statistic of an arc length is fac * (max_i |S_{i+L} - S_i| - 0.5)^2: it is not monotone in |d|, the
// reference's block-pair bound is not an upper bound any more, and which pairs it skips depends on the ORDER in which it
// visits them.  The only way to get its answer is to walk the pairs in its order.  One warp per vector does that:
// lane 0 runs the sequential control flow (prefix sums, pair list, libstdc++'s introsort order of the corner statistics,
// descending visit with the running maximum), all lanes share the innermost loop over the start positions of an arc length.
#pragma once
#include <cuda_runtime.h>

#include "cbs_core.h"

namespace cbsg {

// ---- std::sort of an index array by key, ties ordered as libstdc++'s introsort orders them (median-of-3 quicksort to depth
// 2*log2 n, heap sort fallback, final insertion sort, threshold 16): the reference sorts candidate block pairs with it
// (CBS.cpp:63-66,160) and visits equal keys in whatever order it leaves them. ----------------------------------------------
struct IdxSort {
    int* a;             // the indices being sorted, a[0..len)
    const double* key;  // key[index]
    __device__ bool less(int x, int y) const { return key[x] < key[y]; }
    __device__ void sift(long hole, long len, int value) {
        const long top = hole;
        long child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (less(a[child], a[child - 1])) --child;
            a[hole] = a[child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[hole] = a[child - 1];
            hole = child - 1;
        }
        long parent = (hole - 1) / 2;
        while (hole > top && less(a[parent], value)) {
            a[hole] = a[parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        a[hole] = value;
    }
    __device__ void heap_range(long lo, long hi) {  // heap sort of a[lo..hi)
        int* base = a + lo;
        int* keep = a;
        a = base;
        const long len = hi - lo;
        for (long parent = (len - 2) / 2; len >= 2; --parent) {
            sift(parent, len, a[parent]);
            if (parent == 0) break;
        }
        for (long last = len - 1; last >= 1; --last) {
            const int value = a[last];
            a[last] = a[0];
            sift(0, last, value);
        }
        a = keep;
    }
    __device__ void slide_back(long pos) {  // a[pos] moves left past larger elements (a smaller one is known to stop it)
        const int val = a[pos];
        long nxt = pos - 1;
        while (less(val, a[nxt])) { a[pos] = a[nxt]; pos = nxt; --nxt; }
        a[pos] = val;
    }
    __device__ void insertion(long lo, long hi) {
        for (long i = lo + 1; i < hi; ++i) {
            if (less(a[i], a[lo])) {
                const int val = a[i];
                for (long k = i; k > lo; --k) a[k] = a[k - 1];
                a[lo] = val;
            } else slide_back(i);
        }
    }
    __device__ void run(long len) {
        if (len <= 0) return;
        long lg = 0;
        for (long m = len; m > 1; m >>= 1) ++lg;
        // the quicksort phase, with an explicit stack of (lo, hi, depth): the right part is sorted first, as the recursion does
        long st_lo[64], st_hi[64], st_d[64];
        int sp = 0;
        st_lo[0] = 0; st_hi[0] = len; st_d[0] = 2 * lg; sp = 1;
        while (sp > 0) {
            --sp;
            long lo = st_lo[sp], hi = st_hi[sp], depth = st_d[sp];
            while (hi - lo > 16) {
                if (depth == 0) { heap_range(lo, hi); break; }
                --depth;
                const long mid = lo + (hi - lo) / 2;
                {   // median of a[lo+1], a[mid], a[hi-1] goes to a[lo]
                    const long x = lo + 1, y = mid, z = hi - 1;
                    long pick;
                    if (less(a[x], a[y])) pick = less(a[y], a[z]) ? y : (less(a[x], a[z]) ? z : x);
                    else pick = less(a[x], a[z]) ? x : (less(a[y], a[z]) ? z : y);
                    const int t = a[lo]; a[lo] = a[pick]; a[pick] = t;
                }
                long f = lo + 1, l = hi;
                for (;;) {
                    while (less(a[f], a[lo])) ++f;
                    --l;
                    while (less(a[lo], a[l])) --l;
                    if (!(f < l)) break;
                    const int t = a[f]; a[f] = a[l]; a[l] = t;
                    ++f;
                }
                // recursion on [f, hi) happens BEFORE the loop continues with [lo, f): the order of the two does not change the
                // result (disjoint ranges), so the right part is simply stacked
                if (sp < 64) { st_lo[sp] = f; st_hi[sp] = hi; st_d[sp] = depth; ++sp; }
                hi = f;
            }
        }
        if (len > 16) {
            insertion(0, 16);
            for (long i = 16; i < len; ++i) slide_back(i);
        } else insertion(0, len);
    }
};

// scratch of one vector, carved out of one allocation: doubles first, then ints
__host__ __device__ inline long long bin_scratch_doubles(int n) {
    const long long nb = block_count(n), nb2 = nb * (nb + 1) / 2;
    const long long d = ((long long)n + 1) + 2 * (nb + 1) + 2 * (nb2 + 1);
    const long long i = 3 * (nb + 1) + 4 * (nb2 + 1);
    return d + (i + 1) / 2 + 8;
}

// warp-wide version of `sxmx < absx` with the SMALLEST index winning among equal maxima (CBS.cpp:186-190)
__device__ __forceinline__ void warp_first_max(double& v, int& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// cbs::tmaxo_impl (CBS.cpp:68-227) as written, for `count` vectors of n values laid end to end; one warp per vector
__global__ void __launch_bounds__(32) k_scan_bin(const double* __restrict__ xall, int n, int count, double tss_in, int al0, int ibin,
                                                 double* scratch_all, double* stat_out, int* left_out, int* right_out) {
    const int lane = threadIdx.x;
    for (int v = blockIdx.x; v < count; v += gridDim.x) {
        const double* x = xall + (long long)v * n;
        const double rn = (double)n;
        const int nb = block_count(n);
        const int nb2 = nb * (nb + 1) / 2;
        double* sx = scratch_all + (long long)v * bin_scratch_doubles(n);
        double* bpsmax = sx + n + 1;
        double* bpsmin = bpsmax + nb + 1;
        double* bssbij = bpsmin + nb + 1;
        double* bssijmax = bssbij + nb2 + 1;
        int* bb = (int*)(bssijmax + nb2 + 1);
        int* ibmin = bb + nb + 1;
        int* ibmax = ibmin + nb + 1;
        int* bloci = ibmax + nb + 1;
        int* blocj = bloci + nb2 + 1;
        int* loc = blocj + nb2 + 1;
        int* alen = loc + nb2 + 1;
        double tss = tss_in;
        // :77 block ends, :79-97 prefix sums and block extrema (sequential: lane 0)
        double psmin0 = 0.0, psmax0 = 0.0;
        int ipsmin0 = n, ipsmax0 = n;
        if (lane == 0) {
            bb[0] = 0;
            for (int i = 1; i <= nb; ++i) bb[i] = block_end(n, nb, i);
            sx[0] = 0.0;
            int ilo = 1;
            double psum = 0.0;
            for (int j = 1; j <= nb; ++j) {
                sx[ilo] = psum + x[ilo - 1];
                double psmin = sx[ilo], psmax = sx[ilo];
                int ipsmin = ilo, ipsmax = ilo;
                for (int i = ilo + 1; i <= bb[j]; ++i) {
                    sx[i] = sx[i - 1] + x[i - 1];
                    if (sx[i] < psmin) { psmin = sx[i]; ipsmin = i; }
                    if (sx[i] > psmax) { psmax = sx[i]; ipsmax = i; }
                }
                ibmin[j] = ipsmin; ibmax[j] = ipsmax;
                bpsmin[j] = psmin; bpsmax[j] = psmax;
                if (psmin < psmin0) { psmin0 = psmin; ipsmin0 = ipsmin; }
                if (psmax > psmax0) { psmax0 = psmax; ipsmax0 = ipsmax; }
                psum = sx[bb[j]];
                ilo = bb[j] + 1;
            }
        }
        __syncwarp();
        psmin0 = __shfl_sync(0xffffffffu, psmin0, 0); psmax0 = __shfl_sync(0xffffffffu, psmax0, 0);
        ipsmin0 = __shfl_sync(0xffffffffu, ipsmin0, 0); ipsmax0 = __shfl_sync(0xffffffffu, ipsmax0, 0);
        const double psdiff = psmax0 - psmin0;
        double bssmax = 0.0;
        int tmaxi = min(ipsmax0, ipsmin0), tmaxj = max(ipsmax0, ipsmin0);
        if (psdiff <= 0.0) {  // :101-110
            if (tss <= 0.0001) tss = 1.0;
            bssmax = ibin ? 0.0 / (tss / rn) : 0.0 / ((tss - 0.0) / (rn - 2.0));
            if (lane == 0) { stat_out[v] = bssmax; left_out[v] = tmaxi; right_out[v] = tmaxj; }
            continue;
        }
        {
            const double rj = (double)abs(ipsmax0 - ipsmin0);
            const double rnjov1 = rn / (rj * (rn - rj));
            bssmax = ibin ? rnjov1 * ((psdiff - 0.5) * (psdiff - 0.5)) : rnjov1 * psdiff * psdiff;  // pow(., 2.0) is the exact square
        }
        const double rnov2 = rn / 2.0;
        const int nal0 = n - al0;
        // :119-158 candidate pairs in (i, j) order; the list keeps that order (ballot compaction, 32 pairs at a time)
        int l = 0;
        for (int i = 1; i <= nb; ++i) {
            for (int j0 = i; j0 <= nb; j0 += 32) {
                const int j = j0 + lane;
                bool take = false;
                double e_lim = 0.0, e_bij = 0.0;
                int e_len = 0;
                if (j <= nb) {
                    const int ilo1 = (i == 1) ? 1 : bb[i - 1] + 1, ihi = bb[i];
                    const int jlo = (j == 1) ? 1 : bb[j - 1] + 1, jhi = bb[j];
                    int alenhi = jhi - ilo1;
                    if (alenhi > nal0) alenhi = nal0;
                    int alenlo = (i == j) ? 1 : (jlo - ihi);
                    if (alenlo < al0) alenlo = al0;
                    const double sij1 = fabs(bpsmax[j] - bpsmin[i]);
                    const double sij2 = fabs(bpsmax[i] - bpsmin[j]);
                    const double sijmx0 = fmax(sij1, sij2);
                    const double rjlo = (double)alenlo, rjhi = (double)alenhi;
                    const double rnjov1 = rn / fmin(rjlo * (rn - rjlo), rjhi * (rn - rjhi));
                    const double bsslim = ibin ? rnjov1 * ((sijmx0 - 0.5) * (sijmx0 - 0.5)) : rnjov1 * sijmx0 * sijmx0;
                    if (bssmax <= bsslim) {
                        take = true;
                        e_lim = bsslim;
                        const double sij = (sij1 > sij2) ? sij1 : sij2;
                        e_len = (sij1 > sij2) ? abs(ibmax[j] - ibmin[i]) : abs(ibmin[j] - ibmax[i]);
                        const double rr = (double)e_len;
                        const double fac = rn / (rr * (rn - rr));
                        e_bij = ibin ? fac * ((sij - 0.5) * (sij - 0.5)) : fac * sij * sij;
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, take);
                if (take) {
                    const int pos = l + 1 + __popc(mask & ((1u << lane) - 1u));
                    loc[pos] = pos; bloci[pos] = i; blocj[pos] = j; bssijmax[pos] = e_lim; alen[pos] = e_len; bssbij[pos] = e_bij;
                }
                l += __popc(mask);
            }
        }
        const int nb1 = l;
        __syncwarp();
        if (lane == 0) { IdxSort srt{loc + 1, bssbij}; srt.run(nb1); }  // :160, ascending by corner statistic
        __syncwarp();
        // :162-216 descending visit
        for (int ll = nb1; ll >= 1; --ll) {
            const int k = loc[ll];
            const double bsslim = bssijmax[k];
            if (bssmax > bsslim) continue;
            const int bi = bloci[k], bj = blocj[k];
            int alenmax = alen[k];
            const int ilo1 = (bi == 1) ? 1 : bb[bi - 1] + 1, ihi = bb[bi];
            const int jlo = (bj == 1) ? 1 : bb[bj - 1] + 1, jhi = bb[bj];
            int alenhi = jhi - ilo1;
            if (alenhi > nal0) alenhi = nal0;
            int alenlo = (bi == bj) ? 1 : (jlo - ihi);
            if (alenlo < al0) alenlo = al0;
            const double rjlo = (double)alenlo, rjhi = (double)alenhi;
            if (alenmax > n - alenmax) alenmax = n - alenmax;
            for (int band = 0; band < 2; ++band) {
                int from, to, step;
                if (band == 0) { if (!((rjlo <= rnov2) && (alenlo <= alenmax))) continue; from = alenlo; to = alenmax; step = 1; }
                else {
                    const int amax2 = n - alenmax;
                    if (!((rjhi >= rnov2) && (alenhi >= amax2))) continue;
                    from = alenhi; to = amax2; step = -1;
                }
                for (int i2j = from; step > 0 ? i2j <= to : i2j >= to; i2j += step) {
                    const int ixlo = max(0, jlo - ilo1 - i2j), ixhi = max(0, ihi + i2j - jhi);
                    double sxmx = 0.0;
                    int sxmxi = ilo1;
                    for (int i = ilo1 + ixlo + lane; i <= ihi - ixhi; i += 32) {
                        const double absx = fabs(sx[i + i2j] - sx[i]);
                        if (sxmx < absx) { sxmx = absx; sxmxi = i; }
                    }
                    warp_first_max(sxmx, sxmxi);
                    const double rr = (double)i2j;
                    const double fac = rn / (rr * (rn - rr));
                    const double bijbss = ibin ? fac * ((sxmx - 0.5) * (sxmx - 0.5)) : fac * sxmx * sxmx;
                    if (bijbss > bssmax) { bssmax = bijbss; tmaxi = sxmxi; tmaxj = sxmxi + i2j; }
                }
            }
        }
        if (ibin) {  // :218-224
            if (tss <= 0.0001) tss = 1.0;
            bssmax /= (tss / rn);
        } else {
            if (tss <= bssmax + 0.0001) tss = bssmax + 1.0;
            bssmax /= ((tss - bssmax) / (rn - 2.0));
        }
        if (lane == 0) { stat_out[v] = bssmax; left_out[v] = tmaxi; right_out[v] = tmaxj; }
    }
}

// cbs::btmax (CBS.cpp:363-376): one dependent chain
__global__ void k_btmax(const double* __restrict__ x, int n, double* out) {
    if (threadIdx.x || blockIdx.x) return;
    double sumxi = x[0], ostat = 0.0, di = 1.0;
    const double dn = (double)n;
    for (int i = 2; i <= n - 2; ++i) {
        di += 1.0;
        sumxi += x[i - 1];
        const double b = dn * (sumxi * sumxi) / (di * (dn - di));
        if (ostat < b) ostat = b;
    }
    *out = sqrt(ostat);
}

}  // namespace cbsg
