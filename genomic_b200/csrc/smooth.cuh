// smooth.cuh -- outlier smoothing (DNAcopy smooth.CNA as ported in lib/cbs/smooth.cpp).
//
// Per "group" (one call of cbs::smooth; on the cna segment path one chromosome of one
// sample, cna_segment.hpp:139-140):
//   k_sm_compact  drop non-finite values, keep their indices            (smooth.cpp:135-141)
//   k_sm_diffs    |v[i+1]-v[i]| of the finite subsequence               (smooth.cpp:37-39)
//   (ascending sort of the differences within each group: two radix sorts, by value and then -- stable -- by group;
//    cub::DeviceSegmentedSort for very large calls)
//   k_sm_sd       sum of the n_keep smallest squares in ascending order -> trimmed SD (:36-44,:144-148)
//   k_sm_window   per finite marker: +-k window outlier test, median shrink (smooth.cpp:76-115)
// The window kernel is the HBM-streaming part (24 B per marker); the sort only feeds one
// scalar per group but has to be exact because the sum is taken in sorted order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cbsg {

struct SmoothGroupOut {
    double oSD, sSD;
    int valid;   // thresholds usable (>= 2 finite values, finite non-negative variance)
    int m;       // number of finite values
};

// one CTA per group: stable compaction of finite values
__global__ void __launch_bounds__(256) k_sm_compact(const double* __restrict__ x, const long long* __restrict__ off,
                                                    const int* __restrict__ lab, int n_groups, double* __restrict__ fv,
                                                    int* __restrict__ fidx, int* __restrict__ flab,
                                                    SmoothGroupOut* __restrict__ gout) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long lo = off[g];
        const int n = (int)(off[g + 1] - lo);
        if (threadIdx.x == 0) s_base = 0;
        __syncthreads();
        for (int c0 = 0; c0 < n; c0 += 256) {
            const int i = c0 + threadIdx.x;
            double v = 0.0;
            bool fin = false;
            if (i < n) { v = x[lo + i]; fin = isfinite(v); }
            const unsigned m = __ballot_sync(0xffffffffu, fin);
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            if (lane == 0) s_warp[w] = __popc(m);
            __syncthreads();
            int before = s_base;
            for (int k = 0; k < w; ++k) before += s_warp[k];
            const int pos = before + __popc(m & ((1u << lane) - 1u));
            if (fin) {
                fv[lo + pos] = v;
                fidx[lo + pos] = i;
                if (flab) flab[lo + pos] = lab[lo + i];
            }
            __syncthreads();
            if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 8; ++k) t += s_warp[k]; s_base += t; }
            __syncthreads();
        }
        if (threadIdx.x == 0) { gout[g].m = s_base; gout[g].valid = 0; gout[g].oSD = 0.0; gout[g].sSD = 0.0; }
        __syncthreads();
    }
}

// |differences| of the finite subsequence; slots beyond m-1 of a group are filled with +inf so
// that a segmented sort over the fixed group extents leaves them at the end
// gid (optional): the group of every slot, the second key of the two-pass radix sort (smooth_device)
__global__ void k_sm_diffs(const double* __restrict__ fv, const long long* __restrict__ off, int n_groups,
                           const SmoothGroupOut* __restrict__ gout, double* __restrict__ d, int* __restrict__ gid) {
    for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
        const long long lo = off[g];
        const int n = (int)(off[g + 1] - lo);
        const int m = gout[g].m;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            d[lo + i] = (i + 1 < m) ? fabs(fv[lo + i + 1] - fv[lo + i]) : __longlong_as_double(0x7ff0000000000000LL);
            if (gid) gid[lo + i] = g;
        }
    }
}

// one warp per group: ascending-order sum of squares, exactly as the reference accumulates it (lane 0 runs the
// dependent chain over chunks staged in shared memory, chain_sum_sq in kernels.cuh)
__global__ void __launch_bounds__(32) k_sm_sd(const double* __restrict__ dsorted, const long long* __restrict__ off,
                                              int n_groups, double trim, double inflfact, double outlier_scale,
                                              double smooth_scale, SmoothGroupOut* __restrict__ gout) {
    __shared__ __align__(16) double buf[PREP_CHUNK];
    const int lane = threadIdx.x;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int m = gout[g].m;
        if (m < 2) continue;  // smooth.cpp:142
        double tvar = 0.0;
        const long long keep = llround((1.0 - 2.0 * trim) * (double)(m - 1));  // smooth.cpp:36
        if (keep > 0) {
            double sum = 0.0, ss = 0.0;
            chain_sum_sq(dsorted + off[g], (int)keep, buf, lane, sum, ss);
            tvar = inflfact * (ss / (2.0 * (double)keep));
        }
        if (lane == 0) {
            if (isfinite(tvar) && !(tvar < 0.0)) {  // smooth.cpp:145
                const double sd = sqrt(tvar);
                gout[g].oSD = outlier_scale * sd;
                gout[g].sSD = smooth_scale * sd;
                gout[g].valid = 1;
            }
        }
    }
}

#define SM_MAX_REGION 64

// thread per finite marker
__global__ void __launch_bounds__(256) k_sm_window(const double* __restrict__ fv, const int* __restrict__ fidx,
                                                   const int* __restrict__ flab, const long long* __restrict__ off,
                                                   int n_groups, int region, const SmoothGroupOut* __restrict__ gout,
                                                   double* __restrict__ out) {
    for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
        const SmoothGroupOut go = gout[g];
        if (!go.valid) continue;
        const long long lo = off[g];
        const double* v = fv + lo;
        const int m = go.m;
        const double oSD = go.oSD, sSD = go.sSD;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
            // window clipped to the run of equal labels (smooth.cpp:47-61, :84-85)
            int wlo = i, whi = i;
            if (flab) {
                const int* lb = flab + lo;
                const int mine = lb[i];
                for (int s = 1; s <= region && i - s >= 0 && lb[i - s] == mine; ++s) wlo = i - s;
                for (int s = 1; s <= region && i + s < m && lb[i + s] == mine; ++s) whi = i + s;
            } else {
                wlo = max(0, i - region);
                whi = min(m - 1, i + region);
            }
            const double xi = v[i];
            double above = 100.0 * oSD, below = 100.0 * oSD;
            bool keep = false;
            for (int j = wlo; j <= whi; ++j) {
                if (j == i) continue;
                const double dist = xi - v[j];
                if (fabs(dist) <= oSD) { keep = true; break; }
                if (dist < above) above = dist;
                if (-dist < below) below = -dist;
            }
            double y = xi;
            if (!keep && !((above <= 0.0) && (below <= 0.0))) {
                double buf[2 * SM_MAX_REGION + 1];
                const int cnt = whi - wlo + 1;
                for (int j = 0; j < cnt; ++j) {  // insertion sort, ascending
                    const double val = v[wlo + j];
                    int k = j;
                    while (k > 0 && buf[k - 1] > val) { buf[k] = buf[k - 1]; --k; }
                    buf[k] = val;
                }
                const int h = cnt / 2;
                const double med = (cnt == 2 * h) ? (buf[h - 1] + buf[h]) / 2.0 : buf[h];
                if (above > 0.0) y = med + sSD;
                if (below > 0.0) y = med - sSD;
            }
            out[lo + fidx[lo + i]] = y;
        }
    }
}

// cngpld::summarize_cn (lib/cngpld/summarize.cpp:41-75): thread per output position; `pos_off` delimits the positions of a
// unit, `seg_off` its segments.  Overlap and altered counts and the sum follow the reference's loop order.
__global__ void k_summarize_cn(const unsigned long long* __restrict__ start, const unsigned long long* __restrict__ end,
                               const float* __restrict__ value, const long long* __restrict__ seg_off,
                               const long long* __restrict__ pos_off, int n_units, const unsigned long long* __restrict__ pos,
                               int direction, double cutoff, double* __restrict__ out) {
    const long long total = pos_off[n_units];
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_units;  // unit u with pos_off[u] <= p < pos_off[u+1]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pos_off[mid] <= p) lo = mid; else hi = mid; }
        const unsigned long long q = pos[p];
        long long overlap = 0, altered = 0;
        double sum = 0.0;
        for (long long k = seg_off[lo]; k < seg_off[lo + 1]; ++k) {
            if (start[k] <= q && q <= end[k]) {
                ++overlap;
                const double adj = (double)direction * (double)value[k];
                if (adj > cutoff) { sum += exp(adj); ++altered; }
            }
        }
        out[p] = altered == 0 ? 0.0 : sum / (double)overlap;
    }
}

__global__ void k_count_nonfinite(const double* x, long long n, int* flag) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (!isfinite(x[i])) *flag = 1;
}

}  // namespace cbsg
