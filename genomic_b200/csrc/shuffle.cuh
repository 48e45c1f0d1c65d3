// shuffle.cuh -- xperm (CBS.cpp:487-493) for a whole CTA: an exact, parallel replay of the reference's
// Fisher-Yates shuffle.
//
// The reference runs   for i = n..1:  j = int(u_i * i) + 1;  swap(px[i-1], px[j-1])   with n uniforms in sequence.
// Position i-1 is final after step i, and what it receives is whatever sits at position j-1 at that moment.  A
// position is only ever changed by a step that targets it, so the content of position q-1 just before step i is
//     x[q-1]                     if no earlier step (t > i) has targeted q, else
//     src(t*), t* = min{t > i : j_t = q},  src(t) = content of position t-1 just before step t
// (the swap of step t* parked the old content of position t*-1 there), and src(t) follows the same rule with q = t.
// So with  last[q] = the most recent step that targeted q  (steps that draw j = i swap nothing and are left out):
//     step i reads  old = last[j_i],  writes  last[j_i] = i;     px[i-1] = old ? x[root(old)-1] : x[j_i-1]
//     root(t) = follow t -> last[t] while it is set; every hop goes to a LARGER step, i.e. one that is already done,
//     and whose entry can no longer change (only steps above t target t).
// Steps only interact through last[] of their own target, so a chunk of C consecutive steps is taken at once by the
// CTA: steps with the same target are serialised in step order by a claim table (32-bit atomicMax of
// (epoch, step) on slot hash(target); the largest unresolved step wins its slot, updates last[] and leaves, the
// others try again in the next round of the same chunk).  A round costs two barriers; nearly all steps of a chunk
// finish in the first round.  Root walks and the gather x[.] -> S row happen after the chunk's rounds and overlap the
// next chunk (they only read entries above the chunk).  The result equals the sequential loop bit for bit: same
// uniforms (draw number n-i belongs to step i), same index arithmetic, same final arrangement.
//
// Storage: last[] is the only random-access array, 16 bit per marker in shared memory (segments up to 65535
// markers) or 32 bit in global memory (L2) for longer segments.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cbs_core.h"

namespace cbsg {

struct LastSmem16 {
    unsigned short* a;  // [n+1], 1-based targets
    __device__ __forceinline__ int ld(int q) const { return a[q]; }
    __device__ __forceinline__ void st(int q, int v) const { a[q] = (unsigned short)v; }
    template <int T> __device__ __forceinline__ void clear(int n) const {
        // 32-bit stores; the array starts on a 4-byte boundary
        unsigned* w = (unsigned*)a;
        for (int k = threadIdx.x; k <= n / 2; k += T) w[k] = 0u;
    }
};
struct LastGlobal32 {
    unsigned* a;
    __device__ __forceinline__ int ld(int q) const { return (int)__ldcg(a + q); }
    __device__ __forceinline__ void st(int q, int v) const { __stcg(a + q, (unsigned)v); }
    template <int T> __device__ __forceinline__ void clear(int n) const {
        for (int k = threadIdx.x; k <= n; k += T) __stcg(a + k, 0u);
    }
};

// uniform source of one permutation: draw d (0-based) belongs to step i = n - d
template <bool MT>
struct ShufDraws {
    const uint64_t* win;  // MT: raw (untempered) words of this permutation
    uint32_t k0, k1, permno;  // philox: task key and permutation number (cbs_core.h DrawSrc, stage 0)
    __device__ __forceinline__ uint64_t raw(int d) const {
        if (MT) return __ldg(win + d);
        uint32_t o[4];
        philox4x32_10((uint32_t)d >> 1, permno, 0u, 0u, k0, k1, o);
        return (d & 1) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
    }
    __device__ __forceinline__ uint64_t u64(uint64_t r) const { return MT ? mt_temper(r) : r; }
    // the stream lives in HBM: pull the 128-byte line that holds draw d into L2 ahead of its use
    __device__ __forceinline__ void prefetch(int d) const {
        if (MT) asm volatile("prefetch.global.L2 [%0];" ::"l"(win + d));
    }
};

enum { SHUF_EPOCH_SHIFT = 20, SHUF_EPOCH_MAX = 4000 };

// One permutation by a CTA of T threads, K steps per thread and chunk.
//   claim : [hmask+1] words in shared memory, all below (epoch << 20) on entry; epoch is uniform over the CTA
//   vals  : the values to permute (x of the pending segment), out[i] receives px[i-1] (the S row, 1-based)
//   rdiv  : weighted CBS (wxperm, CBS.cpp:538-547): position i-1 receives y/rw[i-1] unless the step drew j == i
template <int T, int K, class Last, class Draws>
__device__ __forceinline__ void shuffle_cta(Last last, unsigned* claim, int hmask, unsigned& epoch, int n, const Draws src,
                                            const double* __restrict__ vals, const double* __restrict__ rdiv, double* out) {
    const int tid = threadIdx.x;
    last.template clear<T>(n);
    double pend_v[K];
    int pend_i[K];
    uint64_t raw[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        pend_i[k] = 0; pend_v[k] = 0.0;
        const int i = n - k * T - tid;
        raw[k] = (i >= 1) ? src.raw(n - i) : 0ull;
    }
    __syncthreads();
    for (int i0 = n; i0 >= 1; i0 -= K * T) {
        uint64_t nxt[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i2 = i0 - (K + k) * T - tid;
            nxt[k] = (i2 >= 1) ? src.raw(n - i2) : 0ull;
        }
        if (tid < K * T / 16) {  // one thread per 128-byte line of the chunk four ahead
            const int d = n - i0 + 4 * K * T + 16 * tid;
            if (d < n) src.prefetch(d);
        }
        int j[K], lnk[K];
        unsigned un = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = i0 - k * T - tid;
            j[k] = 0; lnk[k] = 0;
            if (i >= 1) {
                j[k] = draw_index(src.u64(raw[k]), i);
                if (j[k] == i) lnk[k] = i;  // no swap; the position keeps what earlier steps parked there
                else un |= 1u << k;
            }
        }
        for (;;) {
            if (epoch >= SHUF_EPOCH_MAX) {  // uniform: the keys would overflow, start over with a clean table
                __syncthreads();
                for (int k = tid; k <= hmask; k += T) claim[k] = 0u;
                epoch = 0;
                __syncthreads();
            }
            ++epoch;
            const unsigned ebase = epoch << SHUF_EPOCH_SHIFT;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (un & (1u << k)) atomicMax(claim + (j[k] & hmask), ebase | (unsigned)(i0 - k * T - tid));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (un & (1u << k)) {
                    const int i = i0 - k * T - tid;
                    if (claim[j[k] & hmask] == (ebase | (unsigned)i)) {
                        lnk[k] = last.ld(j[k]);
                        last.st(j[k], i);
                        un &= ~(1u << k);
                    }
                }
            if (!__syncthreads_or((int)un)) break;
        }
        // the values gathered for the previous chunk have had a whole chunk to arrive
#pragma unroll
        for (int k = 0; k < K; ++k) if (pend_i[k] > 0) out[pend_i[k]] = pend_v[k];
        // root walks of the thread's K steps side by side: the loads of one hop are independent of each other
        int idx[K];
        bool more = false;
#pragma unroll
        for (int k = 0; k < K; ++k) { idx[k] = lnk[k] ? lnk[k] : j[k]; more |= lnk[k] != 0; }
        while (more) {
            more = false;
            int nx[K];
#pragma unroll
            for (int k = 0; k < K; ++k) nx[k] = lnk[k] ? last.ld(idx[k]) : 0;
#pragma unroll
            for (int k = 0; k < K; ++k) { if (nx[k]) { idx[k] = nx[k]; more = true; } else lnk[k] = 0; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = i0 - k * T - tid;
            pend_i[k] = 0;
            if (i >= 1) {
                double v = __ldg(vals + idx[k] - 1);
                if (rdiv && j[k] != i) v = v / __ldg(rdiv + i - 1);
                pend_v[k] = v; pend_i[k] = i;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) raw[k] = nxt[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) if (pend_i[k] > 0) out[pend_i[k]] = pend_v[k];
    __syncthreads();  // last[] is cleared again by the next permutation
}

// The same replay for segments of more than 65535 markers: last[] needs 32 bit per marker (591 KB for the longest
// SNP6 chromosome), more than one SM holds, so a thread-block CLUSTER of R CTAs shares it through distributed shared
// memory.  Targets are dealt round-robin: CTA r owns last[q] of every q with q mod R == r.  Every CTA evaluates ALL
// draws of a chunk (index arithmetic is cheap) but only resolves the steps whose target it owns, so claims and last[]
// updates stay in its own shared memory and need CTA barriers only; one cluster barrier per chunk then makes the
// chunk's entries visible for the root walks, the only remote (DSMEM) reads.
template <int T, int K, int R, class Draws>
__device__ __forceinline__ void shuffle_cluster(unsigned* last, unsigned* claim, int hmask, unsigned& epoch, int n, const Draws src,
                                                const double* __restrict__ vals, const double* __restrict__ rdiv, double* out) {
    namespace cg = cooperative_groups;
    static_assert((R & (R - 1)) == 0, "cluster size must be a power of two");
    constexpr int LR = R == 1 ? 0 : R == 2 ? 1 : R == 4 ? 2 : R == 8 ? 3 : 4;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int tid = threadIdx.x;
    for (int k = tid; k <= (n >> LR) + 1; k += T) last[k] = 0u;
    double pend_v[K];
    int pend_i[K];
    uint64_t raw[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        pend_i[k] = 0; pend_v[k] = 0.0;
        const int i = n - k * T - tid;
        raw[k] = (i >= 1) ? src.raw(n - i) : 0ull;
    }
    __syncthreads();
    for (int i0 = n; i0 >= 1; i0 -= K * T) {
        uint64_t nxt[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i2 = i0 - (K + k) * T - tid;
            nxt[k] = (i2 >= 1) ? src.raw(n - i2) : 0ull;
        }
        if (tid < K * T / 16) {  // one thread per 128-byte line of the chunk four ahead
            const int d = n - i0 + 4 * K * T + 16 * tid;
            if (d < n) src.prefetch(d);
        }
        int j[K], lnk[K];
        unsigned un = 0, mine = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = i0 - k * T - tid;
            j[k] = 0; lnk[k] = 0;
            if (i >= 1) {
                j[k] = draw_index(src.u64(raw[k]), i);
                if ((j[k] & (R - 1)) == rank) {
                    mine |= 1u << k;
                    if (j[k] == i) lnk[k] = i;
                    else un |= 1u << k;
                }
            }
        }
        for (;;) {
            if (epoch >= SHUF_EPOCH_MAX) {
                __syncthreads();
                for (int k = tid; k <= hmask; k += T) claim[k] = 0u;
                epoch = 0;
                __syncthreads();
            }
            ++epoch;
            const unsigned ebase = epoch << SHUF_EPOCH_SHIFT;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (un & (1u << k)) atomicMax(claim + ((j[k] >> LR) & hmask), ebase | (unsigned)(i0 - k * T - tid));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (un & (1u << k)) {
                    const int i = i0 - k * T - tid;
                    if (claim[(j[k] >> LR) & hmask] == (ebase | (unsigned)i)) {
                        lnk[k] = (int)last[j[k] >> LR];
                        last[j[k] >> LR] = (unsigned)i;
                        un &= ~(1u << k);
                    }
                }
            if (!__syncthreads_or((int)un)) break;
        }
        cl.barrier_arrive();  // this CTA's entries of the chunk are final
#pragma unroll
        for (int k = 0; k < K; ++k) if (pend_i[k] > 0) out[pend_i[k]] = pend_v[k];
        cl.barrier_wait();    // ... and so are everybody else's
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = i0 - k * T - tid;
            pend_i[k] = 0;
            if (mine & (1u << k)) {
                int idx = j[k];
                if (lnk[k]) {
                    int r = lnk[k];
                    for (;;) {
                        const unsigned* rl = cl.map_shared_rank(last, r & (R - 1));
                        const int nx = (int)rl[r >> LR];
                        if (!nx) break;
                        r = nx;
                    }
                    idx = r;
                }
                double v = __ldg(vals + idx - 1);
                if (rdiv && j[k] != i) v = v / __ldg(rdiv + i - 1);
                pend_v[k] = v; pend_i[k] = i;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) raw[k] = nxt[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) if (pend_i[k] > 0) out[pend_i[k]] = pend_v[k];
    // the caller's cluster barrier (next work item) keeps last[] alive until every CTA has finished its walks
}

}  // namespace cbsg
