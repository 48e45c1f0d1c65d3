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

struct ChangePointResult {  // CBS.hpp:11-16
    int ncpt = 0;
    std::array<int, 2> icpt{{0, 0}};
    std::array<int, 2> iseg{{0, 0}};
    double ostat = 0.0;
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

// cbs::htmaxp, CBS.hpp:34
inline double htmaxp(const std::vector<double>& px, double tss, int k, int al0, bool ibin) {
    Context& c = default_context();
    double stat = 0.0;
    c.check(cbs_gpu_htmaxp(c.get(), px.data(), (int)px.size(), 1, tss, k, al0, ibin, &stat));
    return stat;
}

// cbs::tailp / cbs::btailp / cbs::btmax, CBS.hpp:29-31
inline double tailp(double b, double delta, int m, int ngrid, double tol) {
    Context& c = default_context();
    double out = 0.0;
    c.check(cbs_gpu_tailp(c.get(), b, delta, m, ngrid, tol, &out));
    return out;
}
inline double btailp(double b, int m, int ng, double tol) {
    Context& c = default_context();
    double out = 0.0;
    c.check(cbs_gpu_btailp(c.get(), b, m, ng, tol, &out));
    return out;
}
inline double btmax(const std::vector<double>& x) {
    Context& c = default_context();
    double out = 0.0;
    c.check(cbs_gpu_btmax(c.get(), x.data(), (int)x.size(), &out));
    return out;
}

// cbs::xperm / cbs::wxperm, CBS.hpp:37-42: rng advances by exactly x.size() draws
inline void xperm(const std::vector<double>& x, std::vector<double>& px, std::mt19937_64& rng) {
    Context& c = default_context();
    const auto state = detail::next312(rng);
    px.resize(x.size());
    if (x.empty()) return;
    c.check(cbs_gpu_xperm(c.get(), x.data(), nullptr, (int)x.size(), state.data(), 0, px.data()));
    rng.discard(x.size());
}
inline void wxperm(const std::vector<double>& x, std::vector<double>& px, const std::vector<double>& rwts, std::mt19937_64& rng) {
    if (x.size() != rwts.size()) throw std::invalid_argument("x and rwts must have same length");
    Context& c = default_context();
    const auto state = detail::next312(rng);
    px.resize(x.size());
    if (x.empty()) return;
    c.check(cbs_gpu_xperm(c.get(), x.data(), rwts.data(), (int)x.size(), state.data(), 0, px.data()));
    rng.discard(x.size());
}

// cbs::tpermp, CBS.hpp:35-36.  px is scratch in the reference; it is left untouched here.
inline double tpermp(int n1, int n2, int n, const double* x, std::vector<double>& px, int nperm, std::mt19937_64& rng) {
    (void)px;
    if (n != n1 + n2) throw std::invalid_argument("tpermp: n must be n1 + n2");
    Context& c = default_context();
    cbs_gpu_params p;
    cbs_gpu_default_params(&p);
    p.nperm = nperm; p.rng_mode = CBS_GPU_RNG_MT19937_64;
    const auto state = detail::next312(rng);
    double pv = 0.0;
    uint64_t draws = 0;
    c.check(cbs_gpu_tpermp(c.get(), x, n1, n2, &p, state.data(), &pv, &draws));
    rng.discard(draws);
    return pv;
}

// cbs::wtmaxo, CBS.hpp:54-58.  cwts is recomputed from wts on the device (the reference's callers derive it from wts).
inline BinarySegmentationResult wtmaxo(const std::vector<double>& x, const std::vector<double>& wts, double tss,
                                       const std::vector<double>& cwts, int al0) {
    (void)cwts;
    if (x.size() != wts.size()) throw std::invalid_argument("x and wts must have same length");
    Context& c = default_context();
    BinarySegmentationResult r;
    c.check(cbs_gpu_wtmaxo(c.get(), x.data(), wts.data(), (int)x.size(), tss, al0, &r.statistic, &r.start, &r.end));
    return r;
}

namespace detail {
inline void no_early_boundary(const std::vector<int>& sbdry, int nperm, const char* who) {
    for (int v : sbdry)
        if (v <= nperm) throw std::runtime_error(std::string(who) + ": a sequential boundary that can stop early is not supported");
}
inline ChangePointResult to_result(const cbs_gpu_split& s) {
    ChangePointResult r;
    r.ncpt = s.ncpt; r.icpt = {{s.icpt0, s.icpt1}}; r.iseg = {{s.iseg0, s.iseg1}}; r.ostat = s.ostat;
    return r;
}
}  // namespace detail

// cbs::fndcpt, CBS.hpp:68-80
inline ChangePointResult fndcpt(const std::vector<double>& x, double tss, int nperm, double cpval, bool ibin, bool hybrid, int al0,
                                int hk, double delta, int ngrid, const std::vector<int>& sbdry, double tol, std::mt19937_64& rng) {
    detail::no_early_boundary(sbdry, nperm, "cbs_gpu::fndcpt");
    Context& c = default_context();
    const cbs_gpu_params p = detail::params(cpval, nperm, hybrid, al0, hk, 0, 0.05, tol, ibin, false, 0.05);
    const auto state = detail::next312(rng);
    cbs_gpu_split s;
    uint64_t draws = 0;
    c.check(cbs_gpu_fndcpt(c.get(), x.data(), (int)x.size(), tss, &p, delta, ngrid, state.data(), &s, &draws));
    rng.discard(draws);
    return detail::to_result(s);
}

// cbs::wfindcpt, CBS.hpp:81-97.  rwts / cwts / delta are derived from wts on the device (see cbs_gpu.h).
inline ChangePointResult wfindcpt(const std::vector<double>& x, double tss, const std::vector<double>& wts,
                                  const std::vector<double>& rwts, const std::vector<double>& cwts, int nperm, double cpval,
                                  bool hybrid, int al0, int hk, double delta, int ngrid, const std::vector<int>& sbdry, double tol,
                                  std::mt19937_64& rng) {
    (void)rwts; (void)cwts; (void)delta;
    if (x.size() != wts.size()) throw std::invalid_argument("x and wts must have same length");
    detail::no_early_boundary(sbdry, nperm, "cbs_gpu::wfindcpt");
    Context& c = default_context();
    const cbs_gpu_params p = detail::params(cpval, nperm, hybrid, al0, hk, 0, 0.05, tol, false, false, 0.05);
    const auto state = detail::next312(rng);
    cbs_gpu_split s;
    uint64_t draws = 0;
    c.check(cbs_gpu_wfindcpt(c.get(), x.data(), wts.data(), (int)x.size(), tss, &p, ngrid, state.data(), &s, &draws));
    rng.discard(draws);
    return detail::to_result(s);
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

// cngpld::summarize_cn (lib/cngpld/summarize.hpp:11-34) on the segments of one sample and chromosome.  The reference takes
// a SegmentedSampleSet, a sample name and a chromosome index and looks the segments up; here the caller passes them
// (`Seg` needs the members start, end and value of cna::Segment<rvalue>, lib/Segment.hpp).
namespace cngpld_gpu {

struct CNSummaryPoint {
    unsigned long pos;
    double value;
};
typedef std::vector<CNSummaryPoint> CNSummary;

template <class Seg>
CNSummary summarize_cn(const std::vector<Seg>& segments, int direction, double cutoff,
                       const std::vector<unsigned long>* positions = nullptr) {
    if (direction != 1 && direction != -1) throw std::invalid_argument("direction must be 1 or -1.");
    if (positions && positions->empty()) return CNSummary();  // (an empty array is not "default positions" to the C entry)
    cbs_gpu::Context& c = cbs_gpu::default_context();
    const size_t n = segments.size();
    std::vector<uint64_t> start(n), end(n);
    std::vector<float> value(n);
    for (size_t i = 0; i < n; ++i) { start[i] = segments[i].start; end[i] = segments[i].end; value[i] = (float)segments[i].value; }
    const int64_t seg_off[2] = {0, (int64_t)n};
    std::vector<uint64_t> pos_in;
    int64_t pos_off[2] = {0, 0};
    if (positions) { pos_in.assign(positions->begin(), positions->end()); pos_off[1] = (int64_t)pos_in.size(); }
    const size_t cap = positions ? pos_in.size() : 2 * n;
    std::vector<uint64_t> out_pos(cap + 1);
    std::vector<double> out_val(cap + 1);
    int64_t out_off[2] = {0, 0};
    c.check(cbs_gpu_summarize_cn(c.get(), seg_off, 1, start.data(), end.data(), value.data(), direction, cutoff,
                                 positions ? pos_off : nullptr, positions ? pos_in.data() : nullptr, out_off, out_pos.data(),
                                 out_val.data()));
    CNSummary out((size_t)out_off[1]);
    for (size_t i = 0; i < out.size(); ++i) { out[i].pos = (unsigned long)out_pos[i]; out[i].value = out_val[i]; }
    return out;
}

}  // namespace cngpld_gpu

#endif
