// cbs_gpu.hpp -- header-only C++ mirror of the reference's lib/cbs call surface on top of the C ABI
// (include/cbs_gpu.h, libcbs_cuda.so).  Same names, argument order, defaults and error behaviour
// as lib/cbs/CBS.hpp:29-128 and lib/cbs/smooth.hpp:8-20, in namespace cbs_gpu instead of cbs, so
// that src/cna_segment.hpp:140-141 (or tests/cbs_test.cpp) switch by changing the namespace.
//
//   * status codes become the exceptions the reference throws: CBS_GPU_ERR_INVALID ->
//     std::invalid_argument (smooth.cpp:17,125-126), CBS_GPU_ERR_OVERFLOW -> std::overflow_error,
//     anything else -> std::runtime_error.  There is no CPU fallback.
//   * `std::mt19937_64& rng` stays in/out: the engine state is handed to the device generator as the
//     next 312 raw words, and the caller's engine is advanced by exactly the draws consumed.
#ifndef CBS_GPU_HPP
#define CBS_GPU_HPP

#include <array>
#include <cstdint>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "cbs_gpu.h"

namespace cbs_gpu {

struct BinarySegmentationResult {  // CBS.hpp:18-22
    double statistic = 0.0;
    int start = 0;
    int end = 0;
};

struct SegmentationResult {  // CBS.hpp:24-27
    std::vector<int> lengths;
    std::vector<double> means;
};

class Context {
public:
    explicit Context(int device = 0) {
        const int ids[1] = {device};
        const int rc = cbs_gpu_create(ids, 1, &ctx_);
        if (rc != CBS_GPU_OK) throw std::runtime_error("cbs_gpu_create failed (status " + std::to_string(rc) + "): no CUDA device or library problem");
    }
    ~Context() { cbs_gpu_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    cbs_gpu_ctx* get() const { return ctx_; }
    void check(int rc) const {
        if (rc == CBS_GPU_OK) return;
        const std::string msg = cbs_gpu_last_error(ctx_);
        if (rc == CBS_GPU_ERR_INVALID) throw std::invalid_argument(msg);
        if (rc == CBS_GPU_ERR_OVERFLOW) throw std::overflow_error(msg);
        throw std::runtime_error("cbs_gpu status " + std::to_string(rc) + ": " + msg);
    }

private:
    cbs_gpu_ctx* ctx_ = nullptr;
};

inline Context& default_context() {
    static Context ctx(0);
    return ctx;
}

namespace detail {
// inverse of the MT19937-64 output tempering
inline uint64_t untemper(uint64_t y) {
    y ^= (y >> 43);
    y ^= (y << 37) & 0xFFF7EEE000000000ULL;
    uint64_t z = y;
    for (int i = 0; i < 4; ++i) z = y ^ ((z << 17) & 0x71D67FFFEDA60000ULL);
    y = z;
    z = y;
    for (int i = 0; i < 3; ++i) z = y ^ ((z >> 29) & 0x5555555555555555ULL);
    return z;
}
// the next 312 raw words the engine will produce (the engine itself is not advanced)
inline std::array<uint64_t, 312> next312(const std::mt19937_64& rng) {
    std::mt19937_64 probe = rng;
    std::array<uint64_t, 312> out;
    for (auto& w : out) w = untemper(probe());
    return out;
}
inline cbs_gpu_params params(double alpha, int nperm, bool hybrid, int min_width, int kmax, int nmin, double eta, double tol,
                             bool ibin, bool undo_prune, double cutoff) {
    cbs_gpu_params p;
    cbs_gpu_default_params(&p);
    p.alpha = alpha; p.nperm = nperm; p.hybrid = hybrid; p.min_width = min_width; p.kmax = kmax; p.nmin = nmin;
    p.eta = eta; p.tol = tol; p.ibin = ibin; p.undo_prune = undo_prune; p.undo_prune_cutoff = cutoff;
    p.do_smooth = 0; p.rng_mode = CBS_GPU_RNG_MT19937_64; p.chain = 1;
    return p;
}
}  // namespace detail

// cbs::smooth, smooth.hpp:8-13
inline std::vector<double> smooth(const std::vector<double>& values, const std::vector<int>& chrom, int smooth_region = 10,
                                  double outlier_sd_scale = 4.0, double smooth_sd_scale = 2.0, double trim = 0.025) {
    if (values.size() != chrom.size()) throw std::invalid_argument("values and chrom must have same length");
    if (smooth_region < 0) throw std::invalid_argument("smooth_region must be non-negative");
    std::vector<double> out(values.size());
    Context& c = default_context();
    c.check(cbs_gpu_smooth(c.get(), values.data(), chrom.data(), (int64_t)values.size(), smooth_region, outlier_sd_scale,
                           smooth_sd_scale, trim, out.data()));
    return out;
}

// cbs::smooth_matrix, smooth.hpp:15-20
inline std::vector<std::vector<double>> smooth_matrix(const std::vector<std::vector<double>>& samples, const std::vector<int>& chrom,
                                                      int smooth_region = 10, double outlier_sd_scale = 4.0,
                                                      double smooth_sd_scale = 2.0, double trim = 0.025) {
    std::vector<std::vector<double>> out;
    out.reserve(samples.size());
    for (const auto& s : samples) out.push_back(smooth(s, chrom, smooth_region, outlier_sd_scale, smooth_sd_scale, trim));
    return out;
}

// cbs::tmaxo / cbs::tmaxp, CBS.hpp:32-33
inline BinarySegmentationResult tmaxo(const std::vector<double>& x, double tss, int al0, bool ibin) {
    Context& c = default_context();
    BinarySegmentationResult r;
    c.check(cbs_gpu_tmaxo(c.get(), x.data(), (int)x.size(), tss, al0, ibin, &r.statistic, &r.start, &r.end));
    return r;
}
inline double tmaxp(const std::vector<double>& px, double tss, int al0, bool ibin) {
    Context& c = default_context();
    double stat = 0.0;
    c.check(cbs_gpu_tmaxp(c.get(), px.data(), (int)px.size(), 1, tss, al0, ibin, &stat));
    return stat;
}

// cbs::segment, CBS.hpp:100-113.  `sbdry` must be the boundary `cna segment` builds
// (src/cna_segment.hpp:130: every entry nperm+1, i.e. the sequential stopping rule disabled).
inline SegmentationResult segment(const std::vector<double>& x, bool ibin, double alpha, int nperm, bool hybrid, int min_width,
                                  int kmax, int nmin, double eta, const std::vector<int>& sbdry, double tol, std::mt19937_64& rng,
                                  bool undo_prune = false, double undo_prune_cutoff = 0.05) {
    for (int v : sbdry)
        if (v <= nperm) throw std::runtime_error("cbs_gpu::segment: a sequential boundary that can stop early is not supported");
    Context& c = default_context();
    const cbs_gpu_params p = detail::params(alpha, nperm, hybrid, min_width, kmax, nmin, eta, tol, ibin, undo_prune, undo_prune_cutoff);
    const auto state = detail::next312(rng);
    SegmentationResult r;
    int cap = (int)x.size() + 1, nseg = 0;
    uint64_t draws = 0;
    r.lengths.resize((size_t)cap);
    r.means.resize((size_t)cap);
    c.check(cbs_gpu_segment(c.get(), x.data(), (int)x.size(), &p, state.data(), cap, r.lengths.data(), r.means.data(), &nseg, &draws));
    r.lengths.resize((size_t)nseg);
    r.means.resize((size_t)nseg);
    rng.discard(draws);
    return r;
}

// cbs::segment_weighted, CBS.hpp:115-128 (CBS.cpp:1026-1099).  Same conventions as segment().
inline SegmentationResult segment_weighted(const std::vector<double>& x, const std::vector<double>& weights, double alpha,
                                           int nperm, bool hybrid, int min_width, int kmax, int nmin, double eta,
                                           const std::vector<int>& sbdry, double tol, std::mt19937_64& rng,
                                           bool undo_prune = false, double undo_prune_cutoff = 0.05) {
    if (x.size() != weights.size()) throw std::invalid_argument("x and weights must have same length");
    for (int v : sbdry)
        if (v <= nperm) throw std::runtime_error("cbs_gpu::segment_weighted: a sequential boundary that can stop early is not supported");
    Context& c = default_context();
    const cbs_gpu_params p = detail::params(alpha, nperm, hybrid, min_width, kmax, nmin, eta, tol, false, undo_prune, undo_prune_cutoff);
    const auto state = detail::next312(rng);
    SegmentationResult r;
    int cap = (int)x.size() + 1, nseg = 0;
    uint64_t draws = 0;
    r.lengths.resize((size_t)cap);
    r.means.resize((size_t)cap);
    c.check(cbs_gpu_segment_weighted(c.get(), x.data(), weights.data(), (int)x.size(), &p, state.data(), cap, r.lengths.data(),
                                     r.means.data(), &nseg, &draws));
    r.lengths.resize((size_t)nseg);
    r.means.resize((size_t)nseg);
    rng.discard(draws);
    return r;
}

}  // namespace cbs_gpu

#endif
