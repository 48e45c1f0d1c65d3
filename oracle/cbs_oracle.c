/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See cbs_oracle.h.
 *
 * Plain-C restatement of the reference algorithm.  Every function cites the
 * reference lines it follows (paths relative to /root/reference).  The floating
 * point operation ORDER is part of the contract: sums are strictly sequential,
 * statistics are evaluated as (fac*s)*s, no FMA contraction (-ffp-contract=off).
 */
#include "cbs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* =========================================================================
 * Random sources
 * ========================================================================= */

/* std::mt19937_64 (ISO C++ [rand.predef]; 312 words, m=156) */
static void mt_refill(orc_rng* r) {
    static const uint64_t UPPER = 0xFFFFFFFF80000000ULL, LOWER = 0x7FFFFFFFULL, MATRIX = 0xB5026F5AA96619E9ULL;
    uint64_t* s = r->mt;
    for (int k = 0; k < 312; ++k) {
        const uint64_t y = (s[k] & UPPER) | (s[(k + 1) % 312] & LOWER);
        s[k] = s[(k + 156) % 312] ^ (y >> 1) ^ ((y & 1ULL) ? MATRIX : 0ULL);
    }
    r->mti = 0;
}

void orc_rng_seed_mt(orc_rng* r, uint64_t seed) {
    memset(r, 0, sizeof(*r));
    r->kind = 0;
    r->mt[0] = seed;
    for (int k = 1; k < 312; ++k)
        r->mt[k] = 6364136223846793005ULL * (r->mt[k - 1] ^ (r->mt[k - 1] >> 62)) + (uint64_t)k;
    r->mti = 312;
}

static uint64_t mt_next(orc_rng* r) {
    if (r->mti >= 312) mt_refill(r);
    uint64_t y = r->mt[r->mti++];
    y ^= (y >> 29) & 0x5555555555555555ULL;
    y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
    y ^= (y << 37) & 0xFFF7EEE000000000ULL;
    y ^= (y >> 43);
    return y;
}

/* Philox4x32-10 (Salmon et al. 2011), as the product's fast mode uses it */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

uint64_t orc_task_key(uint64_t seed, uint64_t unit_id, uint32_t lo, uint32_t hi) {
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL);
    h = mix64(h ^ (unit_id + 0x9E3779B97F4A7C15ULL));
    h = mix64(h ^ (((uint64_t)lo << 32) | (uint64_t)hi));
    return h;
}

void orc_rng_seed_philox(orc_rng* r, uint64_t seed) {
    memset(r, 0, sizeof(*r));
    r->kind = 1;
    r->key0 = (uint32_t)seed;
    r->key1 = (uint32_t)(seed >> 32);
}

void orc_rng_set_task(orc_rng* r, uint64_t seed, uint64_t unit_id, uint32_t lo, uint32_t hi) {
    if (r->kind != 1) return;
    const uint64_t k = orc_task_key(seed, unit_id, lo, hi);
    r->key0 = (uint32_t)k;
    r->key1 = (uint32_t)(k >> 32);
}

void orc_rng_begin(orc_rng* r, uint32_t stage, uint32_t perm) {
    r->stage = stage;
    r->perm = perm;
    r->k = 0;
}

uint64_t orc_rng_u64(orc_rng* r) {
    r->draws++;
    if (r->kind == 0) return mt_next(r);
    const uint32_t ctr[4] = {r->k >> 1, r->perm, r->stage, 0u};
    const uint32_t key[2] = {r->key0, r->key1};
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    const uint64_t v = (r->k & 1u) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
    r->k++;
    return v;
}

/* libstdc++ 13 generate_canonical<double,53> over a 64-bit engine: one draw,
 * double(u64) (round to nearest) / 2^64, clamped below 1 (bits/random.tcc);
 * used by runif01, CBS.cpp:53-55 */
double orc_rng_unif(orc_rng* r) {
    const uint64_t v = orc_rng_u64(r);
    double u = (double)v / 18446744073709551616.0;
    if (u >= 1.0) u = nextafter(1.0, 0.0);
    return u;
}

void orc_rng_discard(orc_rng* r, uint64_t n) {
    if (r->kind == 0) {
        for (uint64_t i = 0; i < n; ++i) (void)mt_next(r);
    }
    r->draws += n;
}

/* =========================================================================
 * libstdc++-style introsort of an index array by key (ascending).
 * The reference orders candidate block pairs with std::sort (CBS.cpp:63-66,160),
 * which is not stable; which of two pairs with EQUAL corner statistic is visited
 * first therefore follows libstdc++'s introsort.  This is that algorithm
 * (median-of-3 quicksort to depth 2*log2, heap sort fallback, insertion finish,
 * threshold 16), restated so ties order identically.
 * ========================================================================= */
typedef struct { const double* key; } idx_cmp;
#define LESS(a, b) (c->key[(a)] < c->key[(b)])

static void unguarded_linear_insert(int* last, const idx_cmp* c) {
    const int val = *last;
    int* next = last - 1;
    while (LESS(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void insertion_sort(int* first, int* last, const idx_cmp* c) {
    if (first == last) return;
    for (int* i = first + 1; i != last; ++i) {
        if (LESS(*i, *first)) {
            const int val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(int));
            *first = val;
        } else unguarded_linear_insert(i, c);
    }
}
static void adjust_heap(int* first, long hole, long len, int value, const idx_cmp* c) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LESS(first[child], first[child - 1])) --child;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    long parent = (hole - 1) / 2;
    while (hole > top && LESS(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void heap_sort(int* first, int* last, const idx_cmp* c) {
    const long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            adjust_heap(first, parent, len, first[parent], c);
            if (parent == 0) break;
            --parent;
        }
    }
    while (last - first > 1) {
        --last;
        const int value = *last;
        *last = *first;
        adjust_heap(first, 0, last - first, value, c);
    }
}
static void median_to_first(int* result, int* a, int* b, int* d, const idx_cmp* c) {
    int* pick;
    if (LESS(*a, *b)) {
        if (LESS(*b, *d)) pick = b;
        else if (LESS(*a, *d)) pick = d;
        else pick = a;
    } else if (LESS(*a, *d)) pick = a;
    else if (LESS(*b, *d)) pick = d;
    else pick = b;
    const int t = *result; *result = *pick; *pick = t;
}
static int* unguarded_partition(int* first, int* last, int* pivot, const idx_cmp* c) {
    for (;;) {
        while (LESS(*first, *pivot)) ++first;
        --last;
        while (LESS(*pivot, *last)) --last;
        if (!(first < last)) return first;
        const int t = *first; *first = *last; *last = t;
        ++first;
    }
}
static void introsort_loop(int* first, int* last, long depth, const idx_cmp* c) {
    while (last - first > 16) {
        if (depth == 0) { heap_sort(first, last, c); return; }
        --depth;
        int* mid = first + (last - first) / 2;
        median_to_first(first, first + 1, mid, last - 1, c);
        int* cut = unguarded_partition(first + 1, last, first, c);
        introsort_loop(cut, last, depth, c);
        last = cut;
    }
}
static void sort_indices_like_libstdcxx(int* first, int* last, const double* key) {
    if (first == last) return;
    idx_cmp cc = {key};
    const idx_cmp* c = &cc;
    long n = last - first, lg = 0;
    while (n > 1) { n >>= 1; ++lg; }
    introsort_loop(first, last, 2 * lg, c);
    if (last - first > 16) {
        insertion_sort(first, first + 16, c);
        for (int* i = first + 16; i != last; ++i) unguarded_linear_insert(i, c);
    } else insertion_sort(first, last, c);
}
#undef LESS

/* =========================================================================
 * Max-t arc scan  (tmaxo_impl, CBS.cpp:68-227)
 * ========================================================================= */
static uint64_t g_arc_evals = 0;
uint64_t orc_arc_evals(void) { return g_arc_evals; }
void orc_arc_evals_reset(void) { g_arc_evals = 0; }

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* CBS.cpp:71 */
static int block_count(int n) { return n >= 50 ? (int)lround(sqrt((double)n)) : 1; }
/* CBS.cpp:77 (1-based, bb[0] unused=0) */
static void block_ends(int n, int nb, int* bb) {
    const double rn = (double)n;
    bb[0] = 0;
    for (int b = 1; b <= nb; ++b) bb[b] = (int)lround(rn * ((double)b / (double)nb));
}

typedef struct {
    int ilo, ihi, jlo, jhi; /* prefix-index ranges of the two blocks (1-based) */
    int lenlo, lenhi;       /* admissible arc lengths, clamped to [al0, n-al0] */
} pair_geo;

/* CBS.cpp:124-131 and :168-175 */
static pair_geo pair_geometry(const int* bb, int bi, int bj, int n, int al0) {
    pair_geo g;
    g.ilo = bb[bi - 1] + 1;
    g.ihi = bb[bi];
    g.jlo = bb[bj - 1] + 1;
    g.jhi = bb[bj];
    g.lenhi = imin(g.jhi - g.ilo, n - al0);
    g.lenlo = (bi == bj) ? 1 : (g.jlo - g.ihi);
    if (g.lenlo < al0) g.lenlo = al0;
    return g;
}

static inline double arc_stat(double rn, int len, double s, int ibin) {
    const double rr = (double)len;
    const double fac = rn / (rr * (rn - rr));
    return ibin ? fac * ((s - 0.5) * (s - 0.5)) : fac * s * s; /* CBS.cpp:191-193 (pow(.,2) folds to a product at -O2) */
}

typedef struct { double best; int bi, bj; } running_max;

/* one arc length of one block pair: CBS.cpp:181-195 / :200-214 */
static void scan_one_length(const double* sx, const pair_geo* g, int len, double rn, int ibin, running_max* m) {
    const int skip_lo = imax(0, g->jlo - g->ilo - len);
    const int skip_hi = imax(0, g->ihi + len - g->jhi);
    double widest = 0.0;
    int where = g->ilo;
    for (int i = g->ilo + skip_lo; i <= g->ihi - skip_hi; ++i) {
        const double d = fabs(sx[i + len] - sx[i]);
        if (widest < d) { widest = d; where = i; }
    }
    g_arc_evals += (uint64_t)imax(0, g->ihi - skip_hi - (g->ilo + skip_lo) + 1);
    const double v = arc_stat(rn, len, widest, ibin);
    if (v > m->best) { m->best = v; m->bi = where; m->bj = where + len; }
}

typedef struct { double stat; int i, j; } arc_result; /* i,j 1-based prefix indices */

static arc_result max_t_scan(const double* x, int n, double tss, int al0, int ibin) {
    const double rn = (double)n;
    const int nb = block_count(n);
    const int npairs = nb * (nb + 1) / 2;
    double* sx = (double*)calloc((size_t)n + 2, sizeof(double));
    int* bb = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
    double* bmin = (double*)malloc(sizeof(double) * (size_t)(nb + 1));
    double* bmax = (double*)malloc(sizeof(double) * (size_t)(nb + 1));
    int* amin = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
    int* amax = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
    double* corner = (double*)malloc(sizeof(double) * (size_t)(npairs + 1));
    double* limit = (double*)malloc(sizeof(double) * (size_t)(npairs + 1));
    int* pi = (int*)malloc(sizeof(int) * (size_t)(npairs + 1));
    int* pj = (int*)malloc(sizeof(int) * (size_t)(npairs + 1));
    int* order = (int*)malloc(sizeof(int) * (size_t)(npairs + 1));
    int* clen = (int*)malloc(sizeof(int) * (size_t)(npairs + 1));
    block_ends(n, nb, bb);

    /* prefix sums + per-block extrema, strictly sequential, first occurrence wins (CBS.cpp:79-97) */
    double run = 0.0, gmin = 0.0, gmax = 0.0;
    int gimin = n, gimax = n;
    for (int b = 1; b <= nb; ++b) {
        const int first = bb[b - 1] + 1;
        run = run + x[first - 1];
        sx[first] = run;
        double lo = run, hi = run;
        int ilo = first, ihi = first;
        for (int i = first + 1; i <= bb[b]; ++i) {
            run = run + x[i - 1];
            sx[i] = run;
            if (run < lo) { lo = run; ilo = i; }
            if (run > hi) { hi = run; ihi = i; }
        }
        bmin[b] = lo; bmax[b] = hi; amin[b] = ilo; amax[b] = ihi;
        if (lo < gmin) { gmin = lo; gimin = ilo; }
        if (hi > gmax) { gmax = hi; gimax = ihi; }
    }

    arc_result out;
    running_max m;
    m.best = 0.0;
    m.bi = imin(gimax, gimin);
    m.bj = imax(gimax, gimin);
    const double spread = gmax - gmin;
    if (spread <= 0.0) { /* CBS.cpp:102-111 */
        if (tss <= 0.0001) tss = 1.0;
        out.stat = ibin ? 0.0 / (tss / rn) : 0.0 / ((tss - 0.0) / (rn - 2.0));
        out.i = m.bi; out.j = m.bj;
        goto done;
    }
    m.best = arc_stat(rn, abs(gimax - gimin), spread, ibin); /* CBS.cpp:113-117 */

    /* candidate block pairs whose bound reaches the current max (CBS.cpp:119-158) */
    int ncand = 0;
    for (int bi = 1; bi <= nb; ++bi) {
        for (int bj = bi; bj <= nb; ++bj) {
            const pair_geo g = pair_geometry(bb, bi, bj, n, al0);
            const double s1 = fabs(bmax[bj] - bmin[bi]);
            const double s2 = fabs(bmax[bi] - bmin[bj]);
            const double smax = s1 > s2 ? s1 : s2; /* std::max(s1,s2) */
            const double rlo = (double)g.lenlo, rhi = (double)g.lenhi;
            const double a = rlo * (rn - rlo), b2 = rhi * (rn - rhi);
            const double fac = rn / (b2 < a ? b2 : a); /* std::min(a,b2) */
            const double bound = ibin ? fac * ((smax - 0.5) * (smax - 0.5)) : fac * smax * smax;
            if (m.best <= bound) {
                ++ncand;
                order[ncand] = ncand;
                pi[ncand] = bi; pj[ncand] = bj;
                limit[ncand] = bound;
                if (s1 > s2) {
                    clen[ncand] = abs(amax[bj] - amin[bi]);
                    corner[ncand] = arc_stat(rn, clen[ncand], s1, ibin);
                } else {
                    clen[ncand] = abs(amin[bj] - amax[bi]);
                    corner[ncand] = arc_stat(rn, clen[ncand], s2, ibin);
                }
            }
        }
    }
    sort_indices_like_libstdcxx(order + 1, order + ncand + 1, corner); /* CBS.cpp:160 */

    /* visit candidates best corner first; scan short arcs up, long arcs down (CBS.cpp:162-216) */
    const double half = rn / 2.0;
    for (int v = ncand; v >= 1; --v) {
        const int k = order[v];
        if (m.best > limit[k]) continue;
        const pair_geo g = pair_geometry(bb, pi[k], pj[k], n, al0);
        int lenmax = clen[k];
        if (lenmax > n - lenmax) lenmax = n - lenmax;
        if (((double)g.lenlo <= half) && (g.lenlo <= lenmax))
            for (int len = g.lenlo; len <= lenmax; ++len) scan_one_length(sx, &g, len, rn, ibin, &m);
        lenmax = n - lenmax;
        if (((double)g.lenhi >= half) && (g.lenhi >= lenmax))
            for (int len = g.lenhi; len >= lenmax; --len) scan_one_length(sx, &g, len, rn, ibin, &m);
    }

    if (ibin) { /* CBS.cpp:218-224 */
        if (tss <= 0.0001) tss = 1.0;
        out.stat = m.best / (tss / rn);
    } else {
        if (tss <= m.best + 0.0001) tss = m.best + 1.0;
        out.stat = m.best / ((tss - m.best) / (rn - 2.0));
    }
    out.i = m.bi; out.j = m.bj;
done:
    free(sx); free(bb); free(bmin); free(bmax); free(amin); free(amax);
    free(corner); free(limit); free(pi); free(pj); free(order); free(clen);
    return out;
}

/* CBS.cpp:378-385 */
orc_tmax orc_tmaxo(const double* x, int n, double tss, int al0, int ibin) {
    const arc_result r = max_t_scan(x, n, tss, al0, ibin);
    orc_tmax o = {r.stat, r.i - 1, r.j - 1};
    return o;
}
double orc_tmaxp(const double* px, int n, double tss, int al0, int ibin) {
    return max_t_scan(px, n, tss, al0, ibin).stat;
}

/* =========================================================================
 * Hybrid pieces: htmaxp (CBS.cpp:387-485), tailp/nu/it1tsq/fpnorm (:14-51, :324-339)
 * ========================================================================= */
static double sq(double v) { return v * v; }

static double h_short_arcs(const double* sx, int from, int to_excl_len, int al0, int k, double rn, double spread_sq,
                           int ibin, double out, int mode, int n, int seam) {
    /* mode 0: arcs inside [from, to_excl_len] (regular, :421-437)
       mode 1: wrap-around arcs i in 1..j against i+n-j (:441-458)
       mode 2: arcs straddling the seam (block end) (:462-476) */
    for (int j = al0; j <= k; ++j) {
        const double rj = (double)j;
        const double fac = rn / (rj * (rn - rj));
        if (fac * spread_sq < out) break;
        double widest = 0.0;
        if (mode == 0) {
            for (int i = from; i <= to_excl_len - j; ++i) { const double d = fabs(sx[i + j] - sx[i]); if (widest < d) widest = d; }
        } else if (mode == 1) {
            const int nmj = n - j;
            for (int i = 1; i <= j; ++i) { const double d = fabs(sx[i + nmj] - sx[i]); if (widest < d) widest = d; }
        } else {
            for (int i = seam + 1 - j; i <= seam; ++i) { const double d = fabs(sx[i + j] - sx[i]); if (widest < d) widest = d; }
        }
        const double v = ibin ? fac * sq(fabs(widest) - 0.5) : fac * widest * widest;
        if (out < v) out = v;
    }
    return out;
}

double orc_htmaxp(const double* px, int n, double tss, int k, int al0, int ibin) {
    const double rn = (double)n;
    const int nb = (int)(rn / (double)k);
    double* sx = (double*)calloc((size_t)n + 2, sizeof(double));
    int* bb = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
    double* bmin = (double*)malloc(sizeof(double) * (size_t)(nb + 1));
    double* bmax = (double*)malloc(sizeof(double) * (size_t)(nb + 1));
    block_ends(n, nb, bb);
    double run = 0.0, out = 0.0;
    for (int b = 1; b <= nb; ++b) {
        const int first = bb[b - 1] + 1;
        run = run + px[first - 1];
        sx[first] = run;
        double lo = run, hi = run;
        int ilo = first, ihi = first;
        for (int i = first + 1; i <= bb[b]; ++i) {
            run = run + px[i - 1];
            sx[i] = run;
            if (run < lo) { lo = run; ilo = i; }
            if (run > hi) { hi = run; ihi = i; }
        }
        bmin[b] = lo; bmax[b] = hi;
        const int d = abs(ilo - ihi);
        if (d <= k && d >= al0) {
            const double rj = (double)d;
            const double fac = rn / (rj * (rn - rj));
            const double v = ibin ? fac * sq(bmax[b] - bmin[b] - 0.5) : fac * sq(bmax[b] - bmin[b]);
            if (out < v) out = v;
        }
    }
#define SPREADSQ(s) (ibin ? sq((s) - 0.5) : sq(s))
    out = h_short_arcs(sx, 1, bb[1], al0, k, rn, SPREADSQ(bmax[1] - bmin[1]), ibin, out, 0, n, 0);
    {
        const double a = fabs(bmax[1] - bmin[nb]), b2 = fabs(bmax[nb] - bmin[1]);
        const double s = a < b2 ? b2 : a;
        out = h_short_arcs(sx, 0, 0, al0, k, rn, SPREADSQ(s), ibin, out, 1, n, 0);
    }
    for (int l = 2; l <= nb; ++l) {
        out = h_short_arcs(sx, bb[l - 1] + 1, bb[l], al0, k, rn, SPREADSQ(bmax[l] - bmin[l]), ibin, out, 0, n, 0);
        const double a = fabs(bmax[l] - bmin[l - 1]), b2 = fabs(bmax[l - 1] - bmin[l]);
        const double s = a < b2 ? b2 : a;
        out = h_short_arcs(sx, 0, 0, al0, k, rn, SPREADSQ(s), ibin, out, 2, n, bb[l - 1]);
    }
#undef SPREADSQ
    free(sx); free(bb); free(bmin); free(bmax);
    if (ibin) {
        if (tss <= 0.0001) tss = 1.0;
        return out / (tss / rn);
    }
    if (tss <= out + 0.0001) tss = out + 1.0;
    return out / ((tss - out) / (rn - 2.0));
}

static double phi_cdf(double x) { return 0.5 * erfc(-x / sqrt(2.0)); } /* fpnorm :14-16 */

static double nu_series(double x, double tol) { /* :18-41 */
    if (x > 0.01) {
        double cur = log(2.0) - 2.0 * log(x);
        double prev = cur;
        int k = 2;
        double dk = 0.0;
        for (int i = 1; i <= k; ++i) {
            dk += 1.0;
            cur -= 2.0 * phi_cdf(-x * sqrt(dk) / 2.0) / dk;
        }
        while (fabs((cur - prev) / cur) > tol) {
            prev = cur;
            for (int i = 1; i <= k; ++i) {
                dk += 1.0;
                cur -= 2.0 * phi_cdf(-x * sqrt(dk) / 2.0) / dk;
            }
            k *= 2;
        }
        return exp(cur);
    }
    return exp(-0.583 * x);
}

static double it1tsq(double x, double a) { /* :43-51 */
    double y = x + a - 0.5;
    double out = (8.0 * y) / (1.0 - 4.0 * y * y) + 2.0 * log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
    y = x - 0.5;
    out -= (8.0 * y) / (1.0 - 4.0 * y * y) + 2.0 * log((1.0 + 2.0 * y) / (1.0 - 2.0 * y));
    return out;
}

double orc_tailp(double b, double delta, int m, int ngrid, double tol) { /* :324-339 */
    const double dincr = (0.5 - delta) / (double)ngrid;
    const double bsqrtm = b / sqrt((double)m);
    double tl = 0.5 - dincr, t = 0.5 - 0.5 * dincr, out = 0.0;
    for (int i = 1; i <= ngrid; ++i) {
        tl += dincr;
        t += dincr;
        const double x = bsqrtm / sqrt(t * (1.0 - t));
        const double nux = nu_series(x, tol);
        out += (nux * nux) * it1tsq(tl, dincr);
    }
    out = 9.973557e-2 * pow(b, 3.0) * exp(-b * b / 2.0) * out;
    return 2.0 * out;
}

/* =========================================================================
 * Permutations
 * ========================================================================= */
/* CBS.cpp:487-493: full Fisher-Yates from the top, n draws (including i==1) */
void orc_xperm(const double* x, int n, double* px, orc_rng* rng) {
    memcpy(px, x, sizeof(double) * (size_t)n);
    for (int i = n; i >= 1; --i) {
        const int j = (int)(orc_rng_unif(rng) * (double)i) + 1;
        const double t = px[i - 1]; px[i - 1] = px[j - 1]; px[j - 1] = t;
    }
}

/* CBS.cpp:495-536.  philox: stage is set by the caller, perm index set here. */
double orc_tpermp(int n1, int n2, int n, const double* x, int nperm, orc_rng* rng, double* px) {
    const double rn1 = (double)n1, rn2 = (double)n2, rn = rn1 + rn2;
    if (n1 == 1 || n2 == 1) return 1.0;
    double sum1 = 0.0, sum2 = 0.0, tss = 0.0;
    for (int i = 0; i < n1; ++i) { sum1 += x[i]; tss += x[i] * x[i]; }
    for (int i = n1; i < n; ++i) { sum2 += x[i]; tss += x[i] * x[i]; }
    const double xbar = (sum1 + sum2) / rn;
    tss -= rn * (xbar * xbar);
    int m1;
    double rm1, ostat, tstat;
    if (n1 <= n2) { m1 = n1; rm1 = rn1; ostat = 0.99999 * fabs(sum1 / rn1 - xbar); tstat = (ostat * ostat) * rn1 * rn / rn2; }
    else          { m1 = n2; rm1 = rn2; ostat = 0.99999 * fabs(sum2 / rn2 - xbar); tstat = (ostat * ostat) * rn2 * rn / rn1; }
    tstat /= ((tss - tstat) / (rn - 2.0));
    if (tstat > 25.0 && m1 >= 10) return 0.0;
    int nrej = 0;
    const uint32_t stage = rng->stage;
    for (int np = 1; np <= nperm; ++np) {
        orc_rng_begin(rng, stage, (uint32_t)(np - 1));
        double acc = 0.0;
        memcpy(px, x, sizeof(double) * (size_t)n);
        for (int i = n; i >= n - m1 + 1; --i) {
            const int j = (int)(orc_rng_unif(rng) * (double)i) + 1;
            const double t = px[i - 1]; px[i - 1] = px[j - 1]; px[j - 1] = t;
            acc += px[i - 1];
        }
        if (ostat <= fabs(acc / rm1 - xbar)) ++nrej;
    }
    return (double)nrej / (double)nperm;
}

/* =========================================================================
 * One split decision (fndcpt, CBS.cpp:830-892)
 * ========================================================================= */
orc_cpt orc_fndcpt(const double* x, int n, double tss, int nperm, double cpval, int ibin, int hybrid, int al0, int hk,
                   double delta, int ngrid, double tol, orc_rng* rng) {
    orc_cpt r;
    memset(&r, 0, sizeof(r));
    r.edge_p[0] = r.edge_p[1] = -1.0;
    double* px = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    const orc_tmax obs = orc_tmaxo(x, n, tss, al0, ibin);
    r.ostat = obs.stat;
    r.iseg[0] = obs.start;
    r.iseg[1] = obs.end;
    const double t1 = sqrt(obs.stat);
    const double thresh = obs.stat * 0.99999;
    if (t1 <= 0.1) { r.exit_code = 1; free(px); return r; }
    const int i1 = obs.start + 1, i2 = obs.end + 1;
    const int arc = imin(i2 - i1, n - i2 + i1);
    if (!(t1 >= 7.0 && arc >= 10)) {
        int nrejc;
        if (hybrid) {
            const double p1 = orc_tailp(t1, delta, n, ngrid, tol);
            if (p1 > cpval) { r.exit_code = 4; free(px); return r; }
            nrejc = (int)((cpval - p1) * (double)nperm);
        } else {
            nrejc = (int)(cpval * (double)nperm);
        }
        /* sbdry is all nperm+1 on this path (cna_segment.hpp:130), so the `np >= sbdry[k-1]`
           break (:855,:865) can never fire before np == nperm+1; only nrej > nrejc exits. */
        for (int np = 1; np <= nperm; ++np) {
            orc_rng_begin(rng, 0u, (uint32_t)(np - 1));
            orc_xperm(x, n, px, rng);
            const double p = hybrid ? orc_htmaxp(px, n, tss, hk, al0, ibin) : orc_tmaxp(px, n, tss, al0, ibin);
            r.perms_run = np;
            if (thresh <= p) ++r.nrej;
            if (r.nrej > nrejc) { r.exit_code = 3; free(px); return r; }
        }
    } else {
        r.exit_code = 2;
    }
    if (i2 == n) {
        r.ncpt = 1; r.icpt[0] = obs.start;
    } else if (i1 == 0) {
        r.ncpt = 1; r.icpt[0] = obs.end;
    } else {
        int n1 = i1, n12 = i2, n2 = n12 - n1;
        rng->stage = 1u;
        double p = orc_tpermp(n1, n2, n12, x, nperm, rng, px);
        r.edge_p[0] = p;
        if (p <= cpval) { r.ncpt = 1; r.icpt[0] = obs.start; }
        n12 = n - i1; n2 = n - i2; n1 = n12 - n2;
        rng->stage = 2u;
        p = orc_tpermp(n1, n2, n12, x + i1, nperm, rng, px);
        r.edge_p[1] = p;
        if (p <= cpval && r.ncpt < 2) { r.icpt[r.ncpt] = obs.end; ++r.ncpt; }
    }
    free(px);
    return r;
}

/* =========================================================================
 * undo.splits = "prune" (prune_segments, CBS.cpp:229-320)
 * ========================================================================= */
static double merged_ssq(const int* lseg, int nseg, const double* segsum, const int* cut, int k) {
    double out = 0.0, s = 0.0;
    int c = 0, from = 0;
    for (int part = 0; part <= k; ++part) {
        const int to = (part < k) ? cut[part] : nseg - 1;
        s = 0.0; c = 0;
        for (int i = from; i <= to; ++i) { s += segsum[i]; c += lseg[i]; }
        out += s * s / (double)c;
        from = to + 1;
    }
    return out;
}
static int advance_combination(int* cut, int r, int nmr) { /* returns "left" */
    int i = r - 1;
    while (i >= 0 && cut[i] == nmr + i) --i;
    if (i < 0) return 0;
    ++cut[i];
    for (int j = i + 1; j < r; ++j) cut[j] = cut[j - 1] + 1;
    return cut[0] == nmr ? 0 : 1;
}
static int prune_lengths(const double* x, int n, int* lseg, int nseg, double pcut) {
    const int ncpt = nseg - 1;
    if (ncpt <= 0) return nseg;
    double ssq = 0.0;
    for (int i = 0; i < n; ++i) ssq += x[i] * x[i];
    double* segsum = (double*)calloc((size_t)nseg, sizeof(double));
    int kk = 0;
    for (int i = 0; i < nseg; ++i) for (int j = 0; j < lseg[i]; ++j) segsum[i] += x[kk++];
    const int k = nseg - 1;
    int* cut = (int*)malloc(sizeof(int) * (size_t)k);
    int* best_prev = (int*)malloc(sizeof(int) * (size_t)k);
    int* best_cur = (int*)malloc(sizeof(int) * (size_t)k);
    for (int i = 0; i < k; ++i) { cut[i] = i; best_prev[i] = i; }
    const double wssqk = ssq - merged_ssq(lseg, nseg, segsum, cut, k);
    int result = -1;
    for (int j = k - 1; j >= 1 && result < 0; --j) {
        const int kmj = k - j;
        for (int i = 0; i < j; ++i) { cut[i] = i; best_cur[i] = i; }
        double wssqj = ssq - merged_ssq(lseg, nseg, segsum, cut, j);
        for (;;) {
            if (!advance_combination(cut, j, kmj)) break;
            const double w = ssq - merged_ssq(lseg, nseg, segsum, cut, j);
            if (w <= wssqj) { wssqj = w; for (int i = 0; i < j; ++i) best_cur[i] = cut[i]; }
        }
        if (wssqj / wssqk > 1.0 + pcut) {
            /* keep the j+1 change points of the previous (finer) level */
            int* cums = (int*)malloc(sizeof(int) * (size_t)nseg);
            int s = 0;
            for (int i = 0; i < nseg; ++i) { s += lseg[i]; cums[i] = s; }
            int* out = (int*)malloc(sizeof(int) * (size_t)(j + 2));
            int prev = 0;
            for (int i = 0; i <= j; ++i) { out[i] = cums[best_prev[i]] - prev; prev = cums[best_prev[i]]; }
            out[j + 1] = n - prev;
            memcpy(lseg, out, sizeof(int) * (size_t)(j + 2));
            result = j + 2;
            free(cums); free(out);
            break;
        }
        for (int i = 0; i < j; ++i) best_prev[i] = best_cur[i];
    }
    if (result < 0) { lseg[0] = n; result = 1; }
    free(segsum); free(cut); free(best_prev); free(best_cur);
    return result;
}

/* =========================================================================
 * Recursive driver (segment, CBS.cpp:959-1024)
 * ========================================================================= */
int orc_segment(const double* x, int n, const orc_seg_opts* o, orc_rng* rng, uint64_t seed, uint64_t unit_id, int cap,
                int* lengths, double* means, orc_split_rec* log, int log_cap, int* log_n) {
    int ends_cap = 64, nends = 2;
    int* ends = (int*)malloc(sizeof(int) * (size_t)ends_cap);
    ends[0] = 0; ends[1] = n;
    int done_cap = 64, ndone = 0;
    int* done = (int*)malloc(sizeof(int) * (size_t)done_cap);
    double* cur = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int nlog = 0;
    while (nends > 1) {
        const int k = nends - 1;
        const int lo = ends[k - 1], hi = ends[k], len = hi - lo;
        orc_cpt z;
        memset(&z, 0, sizeof(z));
        int called = 0;
        if (len >= 2 * o->min_width) {
            const int use_hybrid = o->hybrid && (o->nmin < len);
            const double delta = use_hybrid ? (double)(o->kmax + 1) / (double)len : 0.0;
            int flat = 1;
            for (int i = 0; i < len; ++i) if (!(fabs(x[lo + i] - x[lo]) < 1e-12)) { flat = 0; break; }
            if (!flat) {
                double s = 0.0;
                for (int i = 0; i < len; ++i) s += x[lo + i];
                const double avg = s / (double)len;
                double tss = 0.0;
                for (int i = 0; i < len; ++i) cur[i] = x[lo + i] - avg;
                for (int i = 0; i < len; ++i) tss += cur[i] * cur[i];
                orc_rng_set_task(rng, seed, unit_id, (uint32_t)lo, (uint32_t)hi);
                z = orc_fndcpt(cur, len, tss, o->nperm, o->alpha, o->ibin, use_hybrid, o->min_width, o->kmax, delta,
                               100, o->tol, rng);
                called = 1;
            }
        }
        if (log && nlog < log_cap) {
            orc_split_rec* e = &log[nlog];
            e->lo = lo; e->hi = hi; e->ostat = z.ostat; e->iseg0 = z.iseg[0]; e->iseg1 = z.iseg[1];
            e->ncpt = z.ncpt; e->icpt0 = z.icpt[0]; e->icpt1 = z.icpt[1];
            e->perms_run = z.perms_run; e->nrej = z.nrej; e->exit_code = z.exit_code; e->called = called;
            e->edge_p0 = called ? z.edge_p[0] : -1.0; e->edge_p1 = called ? z.edge_p[1] : -1.0;
        }
        ++nlog;
        if (nends + 2 > ends_cap) { ends_cap *= 2; ends = (int*)realloc(ends, sizeof(int) * (size_t)ends_cap); }
        if (z.ncpt == 0) {
            if (ndone == done_cap) { done_cap *= 2; done = (int*)realloc(done, sizeof(int) * (size_t)done_cap); }
            done[ndone++] = ends[k];
            --nends;
        } else if (z.ncpt == 1) {
            ends[k + 1] = ends[k];
            ends[k] = lo + z.icpt[0] + 1;
            nends += 1;
        } else {
            ends[k + 2] = ends[k];
            ends[k] = lo + z.icpt[0] + 1;
            ends[k + 1] = lo + z.icpt[1] + 1;
            nends += 2;
        }
    }
    if (log_n) *log_n = nlog;
    /* change_loc reversed -> segment lengths (:1006-1012) */
    int nseg = ndone;
    int* lseg = (int*)malloc(sizeof(int) * (size_t)(nseg > 0 ? nseg + 2 : 2));
    int prev = 0;
    for (int s = 0; s < nseg; ++s) { const int e = done[ndone - 1 - s]; lseg[s] = e - prev; prev = e; }
    if (o->undo_prune && nseg > 1) nseg = prune_lengths(x, n, lseg, nseg, o->undo_prune_cutoff);
    int ret;
    if (nseg > cap) ret = -nseg;
    else {
        int ll = 0;
        for (int s = 0; s < nseg; ++s) {
            double acc = 0.0;
            for (int i = ll; i < ll + lseg[s]; ++i) acc += x[i];
            lengths[s] = lseg[s];
            means[s] = acc / (double)lseg[s];
            ll += lseg[s];
        }
        ret = nseg;
    }
    free(ends); free(done); free(cur); free(lseg);
    return ret;
}

/* =========================================================================
 * Smoothing (smooth.cpp)
 * ========================================================================= */
static double norm_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
static double norm_pdf(double x) { return 0.39894228040143267794 * exp(-0.5 * x * x); }

/* standard normal quantile: the routine of oracle/shim/boost/math/distributions/normal.hpp
 * (stand-in for boost::math::quantile, smooth.cpp:19-20), restated in C */
double orc_norm_quantile(double p) {
    if (!(p > 0.0 && p < 1.0)) {
        if (p == 0.0) return -INFINITY;
        if (p == 1.0) return INFINITY;
        return NAN;
    }
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01,  -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00};
    const double plow = 0.02425, phigh = 1.0 - plow;
    double x;
    if (p < plow) {
        const double q = sqrt(-2.0 * log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    } else if (p <= phigh) {
        const double q = p - 0.5, r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
    } else {
        const double q = sqrt(-2.0 * log(1.0 - p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    for (int it = 0; it < 2; ++it) {
        const double e = norm_cdf(x) - p;
        const double u = e / norm_pdf(x);
        x = x - u / (1.0 + 0.5 * x * u);
    }
    return x;
}

/* smooth.cpp:13-31 */
double orc_inflfact(double trim) {
    if (!(trim >= 0.0 && trim < 0.5)) return NAN;
    const double a = orc_norm_quantile(1.0 - trim);
    const int ngrid = 10000;
    const double step = (2.0 * a) / ngrid;
    double sum = 0.0;
    for (int i = 0; i < ngrid; ++i) {
        const double left = -a + i * step;
        const double right = left + step;
        const double x = 0.5 * (left + right);
        sum += x * x * norm_pdf(x) / (1.0 - 2.0 * trim);
    }
    return 1.0 / (sum * step);
}

static int cmp_double(const void* pa, const void* pb) {
    const double a = *(const double*)pa, b = *(const double*)pb;
    return (a > b) - (a < b);
}

/* smooth.cpp:33-45; *bad set when the reference would throw */
static double trimmed_var(const double* v, size_t n, double trim, int* bad) {
    if (n < 2) return 0.0;
    const long long keep = llround((1.0 - 2.0 * trim) * (double)(n - 1));
    if (keep <= 0) return 0.0;
    double* d = (double*)malloc(sizeof(double) * (n - 1));
    for (size_t i = 1; i < n; ++i) d[i - 1] = fabs(v[i] - v[i - 1]);
    qsort(d, n - 1, sizeof(double), cmp_double);
    double ss = 0.0;
    for (long long i = 0; i < keep; ++i) ss += d[i] * d[i];
    free(d);
    /* inflfact validates trim (smooth.cpp:16-18 -> std::invalid_argument); with real Boost,
       quantile(nd, 1.0) (trim == 0) raises std::overflow_error under the default policy */
    if (!(trim >= 0.0 && trim < 0.5)) { *bad = 1; return 0.0; }
    if (trim == 0.0) { *bad = 2; return 0.0; }
    const double f = orc_inflfact(trim);
    return f * (ss / (2.0 * (double)keep));
}

/* smooth.cpp:63-74 */
static double window_median(const double* g, int lo, int hi) {
    double w[64];
    double* buf = (hi - lo + 1 <= 64) ? w : (double*)malloc(sizeof(double) * (size_t)(hi - lo + 1));
    const int m = hi - lo + 1;
    for (int j = 0; j < m; ++j) buf[j] = g[lo + j];
    qsort(buf, (size_t)m, sizeof(double), cmp_double);
    const int h = m / 2;
    const double med = (m == 2 * h) ? (buf[h - 1] + buf[h]) / 2.0 : buf[h];
    if (buf != w) free(buf);
    return med;
}

int orc_smooth(const double* values, const int* chrom, int64_t n, int smooth_region, double outlier_sd_scale,
               double smooth_sd_scale, double trim, double* out) {
    if (smooth_region < 0) return 1; /* smooth.cpp:125-126 (size mismatch cannot be expressed here) */
    for (int64_t i = 0; i < n; ++i) out[i] = values[i];
    int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    double* g = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int* lab = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
        if (isfinite(values[i])) { idx[m] = i; g[m] = values[i]; lab[m] = chrom[i]; ++m; }
    int rc = 0;
    if (m >= 2) {
        int bad = 0;
        const double tvar = trimmed_var(g, (size_t)m, trim, &bad);
        if (bad) rc = bad;
        else if (isfinite(tvar) && !(tvar < 0.0)) {
            const double sd = sqrt(tvar);
            const double oSD = outlier_sd_scale * sd, sSD = smooth_sd_scale * sd;
            /* smooth_lr_kernel (smooth.cpp:76-115) over runs of equal labels (:47-61) */
            int64_t run_lo = 0;
            while (run_lo < m) {
                int64_t run_hi = run_lo;
                while (run_hi + 1 < m && lab[run_hi + 1] == lab[run_lo]) ++run_hi;
                for (int64_t i = run_lo; i <= run_hi; ++i) {
                    const int64_t wlo = (i - smooth_region > run_lo) ? i - smooth_region : run_lo;
                    const int64_t whi = (i + smooth_region < run_hi) ? i + smooth_region : run_hi;
                    double above = 100.0 * oSD, below = 100.0 * oSD;
                    int keep = 0;
                    for (int64_t j = wlo; j <= whi; ++j) {
                        if (j == i) continue;
                        const double dist = g[i] - g[j];
                        if (fabs(dist) <= oSD) { keep = 1; break; }
                        if (dist < above) above = dist;
                        if (-dist < below) below = -dist;
                    }
                    double y = g[i];
                    if (!keep && !((above <= 0.0) && (below <= 0.0))) {
                        const double med = window_median(g, (int)wlo, (int)whi);
                        if (above > 0.0) y = med + sSD;
                        if (below > 0.0) y = med - sSD;
                    }
                    out[idx[i]] = y;
                }
                run_lo = run_hi + 1;
            }
        }
    }
    free(idx); free(g); free(lab);
    return rc;
}

/* =========================================================================
 * Cohort loop (Segment::segment_raw, cna_segment.hpp:127-159)
 * ========================================================================= */
int64_t orc_segment_units(const double* values, const int64_t* unit_off, const int* chrom_label,
                          const uint64_t* unit_ids, int n_units, const orc_cohort_opts* o, int64_t cap, int* seg_count,
                          int* lengths, double* means, uint64_t* draws_out, orc_split_rec* log, int64_t log_cap,
                          int64_t* log_n, int* log_unit) {
    orc_rng rng;
    if (o->rng_kind == 0) orc_rng_seed_mt(&rng, o->seed); else orc_rng_seed_philox(&rng, o->seed);
    int64_t total = 0, nlog = 0;
    for (int u = 0; u < n_units; ++u) {
        const int64_t lo = unit_off[u], hi = unit_off[u + 1];
        seg_count[u] = 0;
        if (draws_out) draws_out[u] = 0;
        if (hi <= lo) continue; /* cna_segment.hpp:138 */
        const int n = (int)(hi - lo);
        if (o->rng_kind == 0 && !o->chain) orc_rng_seed_mt(&rng, o->seed);
        const uint64_t d0 = rng.draws;
        double* x = (double*)malloc(sizeof(double) * (size_t)n);
        if (o->do_smooth) {
            int* lab = (int*)malloc(sizeof(int) * (size_t)n);
            for (int i = 0; i < n; ++i) lab[i] = chrom_label ? chrom_label[u] : 1;
            const int rc = orc_smooth(values + lo, lab, n, o->smooth_region, o->outlier_sd_scale, o->smooth_sd_scale,
                                      o->trim, x);
            free(lab);
            if (rc) { free(x); return -2; }
        } else {
            memcpy(x, values + lo, sizeof(double) * (size_t)n);
        }
        int this_log = 0;
        const int64_t room = cap - total;
        const int k = orc_segment(x, n, &o->seg, &rng, o->seed, unit_ids ? unit_ids[u] : (uint64_t)u,
                                  room > 2147483647 ? 2147483647 : (int)room, lengths + total, means + total,
                                  log ? log + nlog : NULL, (int)((log_cap - nlog) > 2147483647 ? 2147483647 : (log_cap - nlog)),
                                  &this_log);
        free(x);
        if (k < 0) return -1;
        if (log_unit) for (int i = 0; i < this_log && nlog + i < log_cap; ++i) log_unit[nlog + i] = u;
        nlog += this_log;
        seg_count[u] = k;
        total += k;
        if (draws_out) draws_out[u] = rng.draws - d0;
    }
    if (log_n) *log_n = nlog;
    return total;
}

/* weighted CBS (segment_weighted and what it calls): same translation unit, see the file's header */
#include "cbs_oracle_weighted.c"
