// TEST INFRASTRUCTURE (oracle/): C-ABI wrapper around the UNMODIFIED reference
// sources /root/reference/lib/cbs/CBS.cpp and smooth.cpp, which oracle/Makefile
// compiles where they lie (outputs only into oracle/_ref/).  Nothing here is
// product code; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load oracle/_ref/libcbs_ref.so.
//
// The reference exposes free functions over std::vector and std::mt19937_64&
// (lib/cbs/CBS.hpp:29-128, lib/cbs/smooth.hpp:8-20).  This file only adapts
// those to plain pointers, plus a restatement of the cohort loop
// Segment::segment_raw (src/cna_segment.hpp:127-159), because the `cna`
// executable itself cannot be linked in this image (Boost program_options etc.).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <vector>

#include "cbs/CBS.hpp"
#include "cbs/smooth.hpp"

namespace {

// sbdry as `cna segment` builds it (cna_segment.hpp:130): every entry nperm+1, so
// the sequential boundary never triggers.  Only indices below
// nrejc*(nrejc+1)/2 + nrejc + 2 are ever read (CBS.cpp:849-855,859-865), so a short
// vector is equivalent and avoids the int overflow / 20 GB of the CLI expression
// for large nperm.
std::vector<int> make_sbdry(int nperm, double alpha) {
    long long nrejc = static_cast<long long>(alpha * static_cast<double>(nperm));
    if (nrejc < 0) nrejc = 0;
    const long long need = nrejc * (nrejc + 1) / 2 + nrejc + 8;
    return std::vector<int>(static_cast<size_t>(need), nperm + 1);
}

struct RefRng {
    std::mt19937_64 eng;
};

} // namespace

extern "C" {

// ---- RNG handle -----------------------------------------------------------
void* ref_rng_new(uint64_t seed) { return new RefRng{std::mt19937_64(seed)}; }
void ref_rng_free(void* h) { delete static_cast<RefRng*>(h); }
void ref_rng_discard(void* h, uint64_t n) { static_cast<RefRng*>(h)->eng.discard(n); }
uint64_t ref_rng_next_u64(void* h) { return static_cast<RefRng*>(h)->eng(); }
double ref_rng_next_canonical(void* h) {
    return std::generate_canonical<double, 53>(static_cast<RefRng*>(h)->eng);
}
// 1 if the handle's engine equals mt19937_64(seed) advanced by `draws`
int ref_rng_equals(void* h, uint64_t seed, uint64_t draws) {
    std::mt19937_64 e(seed);
    e.discard(draws);
    return e == static_cast<RefRng*>(h)->eng ? 1 : 0;
}

// ---- low level kernels (CBS.hpp:29-40) -----------------------------------------
void ref_tmaxo(const double* x, int n, double tss, int al0, int ibin, double* stat, int* start, int* end) {
    const std::vector<double> v(x, x + n);
    const auto r = cbs::tmaxo(v, tss, al0, ibin != 0);
    *stat = r.statistic;
    *start = r.start;
    *end = r.end;
}

double ref_tmaxp(const double* px, int n, double tss, int al0, int ibin) {
    const std::vector<double> v(px, px + n);
    return cbs::tmaxp(v, tss, al0, ibin != 0);
}

double ref_htmaxp(const double* px, int n, double tss, int k, int al0, int ibin) {
    const std::vector<double> v(px, px + n);
    return cbs::htmaxp(v, tss, k, al0, ibin != 0);
}

double ref_tailp(double b, double delta, int m, int ngrid, double tol) { return cbs::tailp(b, delta, m, ngrid, tol); }

void ref_xperm(const double* x, int n, double* px, void* rng) {
    const std::vector<double> v(x, x + n);
    std::vector<double> p;
    cbs::xperm(v, p, static_cast<RefRng*>(rng)->eng);
    std::memcpy(px, p.data(), sizeof(double) * static_cast<size_t>(n));
}

double ref_tpermp(int n1, int n2, int n, const double* x, int nperm, void* rng) {
    std::vector<double> px;
    return cbs::tpermp(n1, n2, n, x, px, nperm, static_cast<RefRng*>(rng)->eng);
}

// Rejection flags of permutations [perm0, perm0 + nperms) of the max-t loop of cbs::fndcpt (CBS.cpp:860-866) for a
// centred segment x, computed with the reference's OWN cbs::xperm and cbs::tmaxp.  Permutation k of the loop consumes
// draws [start_draw + k*n, start_draw + (k+1)*n) of mt19937_64(seed), so the range is cut over `nthreads` host threads,
// each with its own engine advanced by discard().  flags[k - perm0] = (thresh <= pstat).  For tests of very long
// permutation loops (BASELINE configs[4], 100 000 permutations), which one thread cannot replay in minutes.
void ref_perm_reject_flags(const double* x, int n, double tss, double thresh, int al0, uint64_t seed, uint64_t start_draw,
                           int64_t perm0, int64_t nperms, int nthreads, unsigned char* flags) {
    if (nthreads < 1) nthreads = 1;
    const std::vector<double> v(x, x + n);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t]() {
            const int64_t a = perm0 + nperms * t / nthreads, b = perm0 + nperms * (t + 1) / nthreads;
            if (a >= b) return;
            std::mt19937_64 eng(seed);
            eng.discard(start_draw + static_cast<uint64_t>(a) * static_cast<uint64_t>(n));
            std::vector<double> px;
            for (int64_t k = a; k < b; ++k) {
                cbs::xperm(v, px, eng);
                const double pstat = cbs::tmaxp(px, tss, al0, false);
                flags[k - perm0] = (thresh <= pstat) ? 1 : 0;
            }
        });
    }
    for (auto& th : pool) th.join();
}

// ---- rest of the low-level surface (CBS.hpp:29-98), for the tests of the matching cbs_gpu_* entry points ---------------
double ref_btmax(const double* x, int n) { return cbs::btmax(std::vector<double>(x, x + n)); }
double ref_btailp(double b, int m, int ng, double tol) { return cbs::btailp(b, m, ng, tol); }

// rwts / cwts exactly as cbs::segment_weighted derives them (CBS.cpp:1053-1066)
static void derive_weights(const double* w, int n, std::vector<double>& wts, std::vector<double>& rw, std::vector<double>& cw) {
    wts.assign(w, w + n); rw.resize(n); cw.resize(n);
    double wsum = 0.0, csum = 0.0;
    for (int i = 0; i < n; ++i) { rw[i] = std::sqrt(w[i]); wsum += w[i]; }
    const double cwscale = std::sqrt(wsum);
    for (int i = 0; i < n; ++i) { csum += w[i]; cw[i] = csum / cwscale; }
}
void ref_wtmaxo(const double* x, const double* w, int n, double tss, int al0, double* stat, int* start, int* end) {
    std::vector<double> wts, rw, cw;
    derive_weights(w, n, wts, rw, cw);
    const auto r = cbs::wtmaxo(std::vector<double>(x, x + n), wts, tss, cw, al0);
    *stat = r.statistic; *start = r.start; *end = r.end;
}
void ref_wxperm(const double* x, const double* rwts, int n, double* px, void* rng) {
    std::vector<double> p;
    cbs::wxperm(std::vector<double>(x, x + n), p, std::vector<double>(rwts, rwts + n), static_cast<RefRng*>(rng)->eng);
    std::memcpy(px, p.data(), sizeof(double) * static_cast<size_t>(n));
}
void ref_wfindcpt(const double* x, const double* w, int n, double tss, int nperm, double cpval, int hybrid, int al0, int hk,
                  int ngrid, double tol, void* rng, int* ncpt, int* icpt, int* iseg, double* ostat) {
    std::vector<double> wts, rw, cw;
    derive_weights(w, n, wts, rw, cw);
    const std::vector<int> sbdry = make_sbdry(nperm, cpval);
    const auto r = cbs::wfindcpt(std::vector<double>(x, x + n), tss, wts, rw, cw, nperm, cpval, hybrid != 0, al0, hk, 0.0, ngrid, sbdry,
                                 tol, static_cast<RefRng*>(rng)->eng);
    *ncpt = r.ncpt; icpt[0] = r.icpt[0]; icpt[1] = r.icpt[1]; iseg[0] = r.iseg[0]; iseg[1] = r.iseg[1]; *ostat = r.ostat;
}

// out: ncpt, icpt[2], iseg[2], ostat
void ref_fndcpt(const double* x, int n, double tss, int nperm, double cpval, int ibin, int hybrid, int al0, int hk,
                double delta, int ngrid, double tol, void* rng, int* ncpt, int* icpt, int* iseg, double* ostat) {
    const std::vector<double> v(x, x + n);
    const std::vector<int> sbdry = make_sbdry(nperm, cpval);
    const auto r = cbs::fndcpt(v, tss, nperm, cpval, ibin != 0, hybrid != 0, al0, hk, delta, ngrid, sbdry, tol,
                               static_cast<RefRng*>(rng)->eng);
    *ncpt = r.ncpt;
    icpt[0] = r.icpt[0];
    icpt[1] = r.icpt[1];
    iseg[0] = r.iseg[0];
    iseg[1] = r.iseg[1];
    *ostat = r.ostat;
}

// ---- drivers (CBS.hpp:100-113) -----------------------------------------------
// returns number of segments, or -(needed) if cap too small
int ref_segment(const double* x, int n, int ibin, double alpha, int nperm, int hybrid, int min_width, int kmax,
                int nmin, double eta, double tol, void* rng, int undo_prune, double undo_prune_cutoff, int cap,
                int* lengths, double* means) {
    const std::vector<double> v(x, x + n);
    const std::vector<int> sbdry = make_sbdry(nperm, alpha);
    const auto r = cbs::segment(v, ibin != 0, alpha, nperm, hybrid != 0, min_width, kmax, nmin, eta, sbdry, tol,
                                static_cast<RefRng*>(rng)->eng, undo_prune != 0, undo_prune_cutoff);
    const int k = static_cast<int>(r.lengths.size());
    if (k > cap) return -k;
    for (int i = 0; i < k; ++i) {
        lengths[i] = r.lengths[i];
        means[i] = r.means[i];
    }
    return k;
}

int ref_segment_weighted(const double* x, const double* w, int n, double alpha, int nperm, int hybrid, int min_width,
                         int kmax, int nmin, double eta, double tol, void* rng, int undo_prune,
                         double undo_prune_cutoff, int cap, int* lengths, double* means) {
    const std::vector<double> v(x, x + n), wv(w, w + n);
    const std::vector<int> sbdry = make_sbdry(nperm, alpha);
    const auto r = cbs::segment_weighted(v, wv, alpha, nperm, hybrid != 0, min_width, kmax, nmin, eta, sbdry, tol,
                                         static_cast<RefRng*>(rng)->eng, undo_prune != 0, undo_prune_cutoff);
    const int k = static_cast<int>(r.lengths.size());
    if (k > cap) return -k;
    for (int i = 0; i < k; ++i) {
        lengths[i] = r.lengths[i];
        means[i] = r.means[i];
    }
    return k;
}

// ---- smoothing (smooth.hpp:8-13) ---------------------------------------------
// 0 ok, 1 = std::invalid_argument thrown by the reference, 2 = std::overflow_error (trim == 0)
int ref_smooth(const double* values, const int* chrom, int64_t n, int smooth_region, double outlier_sd_scale,
               double smooth_sd_scale, double trim, double* out) {
    try {
        const std::vector<double> v(values, values + n);
        const std::vector<int> c(chrom, chrom + n);
        const auto r = cbs::smooth(v, c, smooth_region, outlier_sd_scale, smooth_sd_scale, trim);
        std::memcpy(out, r.data(), sizeof(double) * static_cast<size_t>(n));
        return 0;
    } catch (const std::invalid_argument&) {
        return 1;
    } catch (const std::overflow_error&) {
        return 2;
    }
}

// ---- cohort loop: restatement of Segment::segment_raw (cna_segment.hpp:127-159) ----
// Units are (sample, chromosome) vectors laid end to end in `values` (float32-valued
// doubles, widened exactly as cna_segment.hpp:137 does).  chrom_label[u] is the
// 1-based chromosome number used as the constant smoothing label (:139).
//   chain != 0 : ONE mt19937_64(seed) shared serially across all units (:129), i.e.
//                what `cna segment` does;
//   chain == 0 : a fresh mt19937_64(seed) per unit (what tests/cbs_test.cpp:231 etc. do
//                per call); units are then independent and may run on `nthreads`.
// Outputs: seg_count[u]; flat lengths/means in unit order (capacity cap);
// draws are not reported here (the restatement in cbs_oracle.c counts them).
// Returns total number of segments, or -1 if cap is too small, -2 on invalid argument.
int64_t ref_segment_units(const double* values, const int64_t* unit_off, const int* chrom_label, int n_units,
                          int do_smooth, int smooth_region, double outlier_sd_scale, double smooth_sd_scale,
                          double trim, double alpha, int nperm, int hybrid, int min_width, int kmax, int nmin,
                          double eta, double tol, int undo_prune, double undo_prune_cutoff, uint64_t seed, int chain,
                          int nthreads, int64_t cap, int* seg_count, int* lengths, double* means) {
    const std::vector<int> sbdry = make_sbdry(nperm, alpha);
    std::vector<cbs::SegmentationResult> results(static_cast<size_t>(n_units));
    std::atomic<int> bad{0};
    auto run_unit = [&](int u, std::mt19937_64& rng) {
        const int64_t lo = unit_off[u], hi = unit_off[u + 1];
        if (hi <= lo) return;  // empty chromosomes are skipped (:138)
        std::vector<double> x(values + lo, values + hi);
        try {
            if (do_smooth) {
                const std::vector<int> chrom(x.size(), chrom_label[u]);
                x = cbs::smooth(x, chrom, smooth_region, outlier_sd_scale, smooth_sd_scale, trim);
            }
            results[static_cast<size_t>(u)] =
                cbs::segment(x, false, alpha, nperm, hybrid != 0, min_width, kmax, nmin, eta, sbdry, tol, rng,
                             undo_prune != 0, undo_prune_cutoff);
        } catch (const std::exception&) {
            bad = 1;
        }
    };
    if (chain) {
        std::mt19937_64 rng(seed);
        for (int u = 0; u < n_units; ++u) run_unit(u, rng);
    } else {
        if (nthreads < 1) nthreads = 1;
        std::atomic<int> next{0};
        auto worker = [&]() {
            for (;;) {
                const int u = next.fetch_add(1);
                if (u >= n_units) break;
                std::mt19937_64 rng(seed);
                run_unit(u, rng);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
        worker();
        for (auto& th : pool) th.join();
    }
    if (bad) return -2;
    int64_t total = 0;
    for (int u = 0; u < n_units; ++u) {
        const auto& r = results[static_cast<size_t>(u)];
        seg_count[u] = static_cast<int>(r.lengths.size());
        if (total + static_cast<int64_t>(r.lengths.size()) > cap) return -1;
        for (size_t i = 0; i < r.lengths.size(); ++i) {
            lengths[total] = r.lengths[i];
            means[total] = r.means[i];
            ++total;
        }
    }
    return total;
}

} // extern "C"
