// cbs_threads.h -- the thread-per-permutation pieces of the hot path, written as plain
// host/device functions so the same code runs in the CUDA kernels and in the host logic
// tests (tests/emul).
//
//   perm_thread   : xperm (CBS.cpp:487-493) + the prefix-sum / block-extrema pass of
//                   tmaxo_impl (CBS.cpp:79-97) for ONE permutation
//   edge_*_thread : the permutation loop body of tpermp (CBS.cpp:524-534)
#pragma once
#include "cbs_core.h"

namespace cbsg {

// block statistics record of one permutation, `bs_stride(nb)` doubles:
//   [0,nb)        block minima of the prefix sums
//   [nb,2nb)      block maxima
//   [2nb,3nb)     as int[2nb]: argmin (1-based prefix index) per block, then argmax per block
//   [3nb]         global min (starts at 0.0), [3nb+1] global max (starts at 0.0)
//   [3nb+2]       as int[2]: global argmin, argmax (start at n)
//   [3nb+3]       result slot (max statistic before normalisation)
struct BlockStats {
    double* base;
    int nb;
    CBS_HD BlockStats(double* b, int nb_) : base(b), nb(nb_) {}
    CBS_HD double* bmin() const { return base; }
    CBS_HD double* bmax() const { return base + nb; }
    CBS_HD int* amin() const { return (int*)(base + 2 * nb); }
    CBS_HD int* amax() const { return (int*)(base + 2 * nb) + nb; }
    CBS_HD double& gmin() const { return base[3 * nb]; }
    CBS_HD double& gmax() const { return base[3 * nb + 1]; }
    CBS_HD int* gidx() const { return (int*)(base + 3 * nb + 2); }
    CBS_HD double& result() const { return base[3 * nb + 3]; }
};

// Sequential prefix sums with per-block first-occurrence extrema (CBS.cpp:79-97).
// `get(k)` returns element k (0-based) of the (permuted) segment.
template <class Get>
CBS_HD void prefix_and_block_stats(Get get, int n, int nb, const int* bb, double* sx, BlockStats bs) {
    double run = 0.0, g_lo = 0.0, g_hi = 0.0;
    int gi_lo = n, gi_hi = n;
    sx[0] = 0.0;
    for (int b = 1; b <= nb; ++b) {
        const int first = bb[b - 1] + 1, last = bb[b];
        run = run + get(first - 1);
        sx[first] = run;
        double lo = run, hi = run;
        int ilo = first, ihi = first;
        for (int i = first + 1; i <= last; ++i) {
            run = run + get(i - 1);
            sx[i] = run;
            if (run < lo) { lo = run; ilo = i; }
            if (run > hi) { hi = run; ihi = i; }
        }
        bs.bmin()[b - 1] = lo; bs.bmax()[b - 1] = hi;
        bs.amin()[b - 1] = ilo; bs.amax()[b - 1] = ihi;
        if (lo < g_lo) { g_lo = lo; gi_lo = ilo; }
        if (hi > g_hi) { g_hi = hi; gi_hi = ihi; }
    }
    bs.gmin() = g_lo; bs.gmax() = g_hi;
    bs.gidx()[0] = gi_lo; bs.gidx()[1] = gi_hi;
}

// Fisher-Yates from the top (CBS.cpp:489-492) on column p of the [k][P] scratch A,
// four steps at a time: the 8 loads of a group are issued together and the aliasing
// between the steps of the group is resolved in registers, so a thread waits for one
// memory round trip per four steps instead of per step.  Bit-identical to the
// sequential loop (checked against the oracle in tests/emul).
CBS_HD void fy_shuffle_column(double* A, long long P, long long p, const double* src_vals, int n, DrawSrc& draws) {
    for (int k = 0; k < n; ++k) A[(long long)k * P + p] = src_vals[k];
    int i = n;
    uint32_t kdraw = 0;
    for (; i >= 4; i -= 4) {
        int j[4];
        double R[4], T[4], W[4], F[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) j[s] = draw_index(draws.u64(kdraw + s), i - s);
        kdraw += 4;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            R[s] = A[(long long)(i - s - 1) * P + p];
            T[s] = A[(long long)(j[s] - 1) * P + p];
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            double ci = R[s], cj = T[s];
#pragma unroll
            for (int m = 0; m < s; ++m) {
                if (j[m] == i - s) ci = W[m];
                if (j[m] == j[s]) cj = W[m];
            }
            if (j[s] == i - s) cj = ci;
            W[s] = ci;  // goes to position j[s]
            F[s] = cj;  // final value of row i-s
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            A[(long long)(i - s - 1) * P + p] = F[s];
            A[(long long)(j[s] - 1) * P + p] = W[s];
        }
    }
    for (; i >= 1; --i) {
        const int jj = draw_index(draws.u64(kdraw++), i);
        const double a = A[(long long)(i - 1) * P + p], b = A[(long long)(jj - 1) * P + p];
        A[(long long)(i - 1) * P + p] = b;
        A[(long long)(jj - 1) * P + p] = a;
    }
}

struct ColumnGet {
    const double* A; long long P, p;
    CBS_HD double operator()(int k) const { return A[(long long)k * P + p]; }
};
struct PlainGet {
    const double* v;
    CBS_HD double operator()(int k) const { return v[k]; }
};

// one permutation of the max-t test: shuffle + prefix sums + block stats
CBS_HD void perm_thread(const Dev& D, const Task& t, int P, int p) {
    const int n = t.n, nb = t.nb;
    const long long base = D.unit_off[t.unit] + t.lo;
    const double* cur = D.cur + base;
    const int* bb = D.bbtab + base;
    double* A = D.arena + t.off_A;
    double* sx = D.arena + t.off_sx + (long long)p * Sched::sx_stride(n);
    BlockStats bs(D.arena + t.off_bs + (long long)p * Sched::bs_stride(nb), nb);
    DrawSrc src;
    if (D.prm.rng_mode == RNG_MT) src.init_mt(draw_window(D, t.off_draw + (long long)p * n));
    else src.init_philox(t.key, 0u, (uint32_t)(t.perms_done + p));
    fy_shuffle_column(A, P, p, cur, n, src);
    ColumnGet g{A, P, p};
    prefix_and_block_stats(g, n, nb, bb, sx, bs);
}

// ---- edge tests -----------------------------------------------------------------------
// tpermp set-up (CBS.cpp:496-522), strictly sequential sums
CBS_HD void edgeprep_seq(Dev& D, Task& t, int s) {
    const int n1 = t.e_n1[s], n2 = t.e_n2[s], n = n1 + n2;
    const double* x = D.cur + D.unit_off[t.unit] + t.lo + t.e_off[s];
    t.e_nrej[s] = 0;
    if (n1 == 1 || n2 == 1) { t.e_status[s] = 1; t.e_m1[s] = 0; return; }
    const double rn1 = (double)n1, rn2 = (double)n2, rn = rn1 + rn2;
    double sum1 = 0.0, sum2 = 0.0, tss = 0.0;
    for (int i = 0; i < n1; ++i) { sum1 += x[i]; tss += x[i] * x[i]; }
    for (int i = n1; i < n; ++i) { sum2 += x[i]; tss += x[i] * x[i]; }
    const double xbar = (sum1 + sum2) / rn;
    tss -= rn * (xbar * xbar);
    int m1; double rm1, ostat, tstat;
    if (n1 <= n2) { m1 = n1; rm1 = rn1; ostat = 0.99999 * fabs(sum1 / rn1 - xbar); tstat = (ostat * ostat) * rn1 * rn / rn2; }
    else          { m1 = n2; rm1 = rn2; ostat = 0.99999 * fabs(sum2 / rn2 - xbar); tstat = (ostat * ostat) * rn2 * rn / rn1; }
    tstat /= ((tss - tstat) / (rn - 2.0));
    t.e_m1[s] = m1; t.e_rm1[s] = rm1; t.e_ostat[s] = ostat; t.e_xbar[s] = xbar;
    t.e_status[s] = (tstat > 25.0 && m1 >= 10) ? 2 : 0;
}
// the scalar tail of edgeprep, shared with the warp version
CBS_HD void edgeprep_finish(Task& t, int s, double sum1, double sum2, double tss) {
    const int n1 = t.e_n1[s], n2 = t.e_n2[s];
    const double rn1 = (double)n1, rn2 = (double)n2, rn = rn1 + rn2;
    const double xbar = (sum1 + sum2) / rn;
    tss -= rn * (xbar * xbar);
    int m1; double rm1, ostat, tstat;
    if (n1 <= n2) { m1 = n1; rm1 = rn1; ostat = 0.99999 * fabs(sum1 / rn1 - xbar); tstat = (ostat * ostat) * rn1 * rn / rn2; }
    else          { m1 = n2; rm1 = rn2; ostat = 0.99999 * fabs(sum2 / rn2 - xbar); tstat = (ostat * ostat) * rn2 * rn / rn1; }
    tstat /= ((tss - tstat) / (rn - 2.0));
    t.e_m1[s] = m1; t.e_rm1[s] = rm1; t.e_ostat[s] = ostat; t.e_xbar[s] = xbar;
    t.e_status[s] = (tstat > 25.0 && m1 >= 10) ? 2 : 0;
}

CBS_HD void edge_draw_src(const Dev& D, const Task& t, const EdgeItem& e, int r /*perm index inside the batch*/, DrawSrc& src) {
    if (D.prm.rng_mode == RNG_MT) src.init_mt(draw_window(D, e.off_draw + (long long)r * t.e_m1[e.side]));
    else src.init_philox(t.key, (uint32_t)(1 + e.side), (uint32_t)(e.perm0 + r));
}

// m1 <= 64: the partial shuffle touches at most 2*m1 positions, kept as an override list in
// local memory; no scratch copy of the segment.  Returns 1 if the permutation rejects.
CBS_HD int edge_sparse_thread(const Dev& D, const Task& t, const EdgeItem& e, int r) {
    const int s = e.side, m1 = t.e_m1[s], n = t.e_n1[s] + t.e_n2[s];
    const double* x = D.cur + D.unit_off[t.unit] + t.lo + t.e_off[s];
    DrawSrc src;
    edge_draw_src(D, t, e, r, src);
    int okey[64];
    double oval[64];
    int cnt = 0;
    double acc = 0.0;
    uint32_t kd = 0;
    for (int i = n; i >= n - m1 + 1; --i) {
        const int j = draw_index(src.u64(kd++), i);
        double vi = x[i - 1];
        for (int q = 0; q < cnt; ++q) if (okey[q] == i - 1) vi = oval[q];
        double vj;
        int slot = -1;
        if (j == i) vj = vi;
        else {
            vj = x[j - 1];
            for (int q = 0; q < cnt; ++q) if (okey[q] == j - 1) { vj = oval[q]; slot = q; }
            if (slot < 0) { slot = cnt++; okey[slot] = j - 1; }
            oval[slot] = vi;
        }
        acc += vj;
    }
    const double pstat = fabs(acc / t.e_rm1[s] - t.e_xbar[s]);
    return t.e_ostat[s] <= pstat ? 1 : 0;
}

// general case: column c of a [k][cols] scratch copy of the segment; the thread runs its Q
// permutations one after another and undoes each partial shuffle by replaying it backwards.
CBS_HD int edge_general_thread(const Dev& D, const Task& t, const EdgeItem& e, int c) {
    const int s = e.side, m1 = t.e_m1[s], n = t.e_n1[s] + t.e_n2[s];
    const double* x = D.cur + D.unit_off[t.unit] + t.lo + t.e_off[s];
    double* A = D.arena + e.off_scratch;
    const long long C = e.cols;
    for (int k = 0; k < n; ++k) A[(long long)k * C + c] = x[k];
    int rejections = 0;
    for (int q = 0; q < e.Q; ++q) {
        const int r = c + q * e.cols;
        if (r >= e.P) break;
        DrawSrc src;
        edge_draw_src(D, t, e, r, src);
        double acc = 0.0;
        uint32_t kd = 0;
        for (int i = n; i >= n - m1 + 1; --i) {
            const int j = draw_index(src.u64(kd++), i);
            const double a = A[(long long)(i - 1) * C + c], b = A[(long long)(j - 1) * C + c];
            A[(long long)(i - 1) * C + c] = b;
            A[(long long)(j - 1) * C + c] = a;
            acc += b;
        }
        const double pstat = fabs(acc / t.e_rm1[s] - t.e_xbar[s]);
        if (t.e_ostat[s] <= pstat) ++rejections;
        if (q + 1 < e.Q && r + e.cols < e.P) {
            for (int i = n - m1 + 1; i <= n; ++i) {  // undo, last swap first
                const int j = draw_index(src.u64((uint32_t)(n - i)), i);
                const double a = A[(long long)(i - 1) * C + c], b = A[(long long)(j - 1) * C + c];
                A[(long long)(i - 1) * C + c] = b;
                A[(long long)(j - 1) * C + c] = a;
            }
        }
    }
    return rejections;
}

// ordered rejection count of one batch (CBS.cpp:863-864): index of the permutation that
// makes nrej exceed nrejc, or -1
CBS_HD void count_item_seq(Dev& D, const PermItem& it) {
    Task& t = D.tasks[it.task];
    if (it.obs) return;
    const int* rej = D.rej + t.off_rej;
    int nrej = t.nrej, hit = -1, inb = 0;
    for (int p = 0; p < it.P; ++p) {
        if (rej[p]) { ++nrej; ++inb; }
        if (nrej > t.nrejc) { hit = p; break; }
    }
    t.cnt_exit = hit; t.cnt_nrej = inb;
}

// MT19937-64 raw stream W[0..) following the chain's cursor: W[w] = hist[w] for w < 312 and
// W[w] = W[w-156] ^ twist(W[w-312], W[w-311]) beyond (the engine's recurrence).
// Commit the `commit_d` words consumed from last round's window (hist <- W_prev[d..d+312)),
// then write this round's window W[0 .. need_len+312).
CBS_HD void mt_generate_seq(Chain& ch, const uint64_t* prev_arena, uint64_t* cur_arena) {
    const uint64_t d = ch.commit_d;
    if (d) for (int u = 0; u < 312; ++u) ch.hist[u] = prev_arena[ch.prev_off + (long long)d + u];
    if (ch.need_len == 0) return;
    uint64_t* out = cur_arena + ch.need_off;
    for (int u = 0; u < 312; ++u) out[u] = ch.hist[u];
    for (uint64_t w = 312; w < ch.need_len + 312; ++w) out[w] = mt_twist(out[w - 312], out[w - 311], out[w - 156]);
}

// shared stream: extend W[0..stream_len) to at least stream_target words (sequential form)
CBS_HD void mt_extend_stream_seq(Dev& D) {
    long long len = D.stream_len;
    while (len < D.stream_target) { stream_put(D, len, mt_twist(stream_get(D, len - 312), stream_get(D, len - 311), stream_get(D, len - 156))); ++len; }
    D.stream_len = len;
}

// the first 312 raw words a freshly seeded std::mt19937_64(seed) will output (untempered)
CBS_HD void mt_seed_next312(uint64_t seed, uint64_t* next) {
    uint64_t st[312];
    st[0] = seed;
    for (int k = 1; k < 312; ++k) st[k] = 6364136223846793005ULL * (st[k - 1] ^ (st[k - 1] >> 62)) + (uint64_t)k;
    for (int k = 0; k < 312; ++k) {
        const uint64_t a = st[k];
        const uint64_t b = (k + 1 < 312) ? st[k + 1] : next[0];
        const uint64_t m = (k + 156 < 312) ? st[k + 156] : next[k + 156 - 312];
        next[k] = mt_twist(a, b, m);
    }
}
// inverse of mt_temper: recovers the raw word from an engine output
CBS_HD uint64_t mt_untemper(uint64_t y) {
    y ^= (y >> 43);
    y ^= (y << 37) & 0xFFF7EEE000000000ULL;
    // y ^= (y << 17) & mask : iterate to undo
    uint64_t z = y;
    for (int i = 0; i < 4; ++i) z = y ^ ((z << 17) & 0x71D67FFFEDA60000ULL);
    y = z;
    z = y;
    for (int i = 0; i < 3; ++i) z = y ^ ((z >> 29) & 0x5555555555555555ULL);
    return z;
}

}  // namespace cbsg
