// GPU test of the header-only shim (genomic_b200/host/cbs_gpu.hpp): the reference's own unit tests
// (tests/cbs_test.cpp:154-203, :287-330; tests/smooth_test.cpp) with namespace cbs -> cbs_gpu, the low-level surface, plus a
// noisy vector checked against the oracle (liboracle.so) including the in/out engine state.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../../genomic_b200/host/cbs_gpu.hpp"
#include "../../oracle/cbs_oracle.h"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL line %d: %s\n", __LINE__, #c); ++fails; } } while (0)

int main() {
    std::vector<double> x;
    for (int i = 0; i < 20; ++i) x.push_back(0.0);
    for (int i = 0; i < 20; ++i) x.push_back(1.5);
    for (int i = 0; i < 20; ++i) x.push_back(0.0);
    double sumx = 0, sumsq = 0;
    for (double v : x) { sumx += v; sumsq += v * v; }
    const double tss = sumsq - (sumx * sumx) / x.size();
    {   // Unweighted_tmaxo_Matches_FortranRawStatisticBehavior
        const auto obs = cbs_gpu::tmaxo(x, tss, 2, false);
        CHECK(obs.start == 0); CHECK(obs.end == 58); CHECK(obs.statistic > 1000.0);
    }
    {   // the same test's ibin = true half (tests/cbs_test.cpp:171-175): [0,58], statistic > 10
        const auto obs = cbs_gpu::tmaxo(x, tss, 2, true);
        CHECK(obs.start == 0); CHECK(obs.end == 58); CHECK(obs.statistic > 10.0);
        CHECK(std::fabs(obs.statistic - 900.259) < 1e-3);  // what the compiled reference returns
    }
    {   // Weighted_wtmaxo_Matches_FortranRawStatisticBehavior (tests/cbs_test.cpp:179-203)
        std::vector<double> xw, ww;
        for (int i = 0; i < 15; ++i) { xw.push_back(0.0); ww.push_back(1.0); }
        for (int i = 0; i < 15; ++i) { xw.push_back(2.0); ww.push_back(0.5); }
        for (int i = 0; i < 15; ++i) { xw.push_back(-1.5); ww.push_back(2.0); }
        for (int i = 0; i < 15; ++i) { xw.push_back(0.0); ww.push_back(1.0); }
        double sw = 0, swx = 0, swxx = 0;
        for (size_t i = 0; i < xw.size(); ++i) { sw += ww[i]; swx += ww[i] * xw[i]; swxx += ww[i] * xw[i] * xw[i]; }
        const double wtss = swxx - (swx * swx) / sw;
        std::vector<double> cw(xw.size());
        double cs = 0.0;
        for (size_t i = 0; i < xw.size(); ++i) { cs += ww[i]; cw[i] = cs / std::sqrt(sw); }
        const auto o2 = cbs_gpu::wtmaxo(xw, ww, wtss, cw, 2);
        CHECK(o2.start == 0); CHECK(o2.end == 58);
        const auto o3 = cbs_gpu::wtmaxo(xw, ww, wtss, cw, 3);
        CHECK(o3.start == 0); CHECK(o3.end == 57);
    }
    {   // low-level surface with the engine in/out (CBS.hpp:35-37,68-80): xperm, tpermp, fndcpt against the oracle restatement
        std::mt19937_64 g(21);
        std::normal_distribution<double> nz(0.0, 0.2);
        std::vector<double> y(1200);
        double s = 0.0;
        for (size_t i = 0; i < y.size(); ++i) { y[i] = (double)(float)(nz(g) + ((i >= 500 && i < 640) ? 0.12 : 0.0)); }
        for (double v : y) s += v;
        const double avg = s / (double)y.size();
        double ytss = 0.0;
        for (auto& v : y) v -= avg;
        for (double v : y) ytss += v * v;
        std::mt19937_64 rng(5);
        orc_rng orng;
        orc_rng_seed_mt(&orng, 5);
        std::vector<double> px, want(y.size());
        cbs_gpu::xperm(y, px, rng);
        orc_xperm(y.data(), (int)y.size(), want.data(), &orng);
        for (size_t i = 0; i < y.size(); ++i) CHECK(px[i] == want[i]);
        std::vector<double> scratch(y.size());
        const double p1 = cbs_gpu::tpermp(500, 140, 640, y.data(), px, 300, rng);
        CHECK(p1 == orc_tpermp(500, 140, 640, y.data(), 300, &orng, scratch.data()));
        const int nperm = 400;
        std::vector<int> sbdry(2000, nperm + 1);
        const auto cp = cbs_gpu::fndcpt(y, ytss, nperm, 0.05, false, false, 2, 25, 0.0, 100, sbdry, 1e-6, rng);
        const orc_cpt oc = orc_fndcpt(y.data(), (int)y.size(), ytss, nperm, 0.05, 0, 0, 2, 25, 0.0, 100, 1e-6, &orng);
        CHECK(cp.ncpt == oc.ncpt); CHECK(cp.iseg[0] == oc.iseg[0] && cp.iseg[1] == oc.iseg[1]); CHECK(cp.ostat == oc.ostat);
        if (oc.ncpt >= 1) CHECK(cp.icpt[0] == oc.icpt[0]);
        if (oc.ncpt == 2) CHECK(cp.icpt[1] == oc.icpt[1]);
        for (int i = 0; i < 5; ++i) CHECK(rng() == orc_rng_u64(&orng));  // three calls later both engines agree
    }
    {   // Unweighted_SegmentDriver_MatchesDNAcopy_SimpleCase (tests/cbs_test.cpp:287-306)
        // all four parameter rows of the reference's fixture table: perm, perm_alt, hybrid, hybrid_alt
        const struct { double alpha; int nperm; bool hybrid; int mw; } cases[] = {{0.01, 200, false, 2}, {0.05, 100, false, 3},
                                                                                 {0.01, 200, true, 2}, {0.05, 100, true, 3}};
        for (const auto& tc : cases) {
            std::mt19937_64 rng(1);
            std::vector<int> sbdry((tc.nperm + 1) * (tc.nperm + 2) / 2 + 2, tc.nperm + 1);
            const auto seg = cbs_gpu::segment(x, false, tc.alpha, tc.nperm, tc.hybrid, tc.mw, 25, 200, 0.05, sbdry, 1e-6, rng, false, 0.05);
            CHECK(seg.lengths.size() == 3);
            if (seg.lengths.size() == 3) {
                CHECK(seg.lengths[0] == 20 && seg.lengths[1] == 20 && seg.lengths[2] == 20);
                CHECK(std::fabs(seg.means[0]) < 1e-9 && std::fabs(seg.means[1] - 1.5) < 1e-9 && std::fabs(seg.means[2]) < 1e-9);
            }
        }
    }
    {   // noisy vector, engine used before and after the call
        std::mt19937_64 g(99);
        std::normal_distribution<double> nz(0.0, 0.2);
        std::vector<double> y(900);
        for (size_t i = 0; i < y.size(); ++i) y[i] = (double)(float)(nz(g) + ((i >= 300 && i < 420) ? 0.5 : 0.0));
        std::mt19937_64 rng(7);
        rng.discard(1000);
        orc_rng orng;
        orc_rng_seed_mt(&orng, 7);
        orc_rng_discard(&orng, 1000);
        const int nperm = 500;
        std::vector<int> sbdry(2000, nperm + 1);
        const auto seg = cbs_gpu::segment(y, false, 0.01, nperm, false, 2, 25, 200, 0.05, sbdry, 1e-6, rng);
        orc_seg_opts o = {0, 0.01, nperm, 0, 2, 25, 200, 0.05, 1e-6, 0, 0.05};
        std::vector<int> len(y.size());
        std::vector<double> mean(y.size());
        const int k = orc_segment(y.data(), (int)y.size(), &o, &orng, 7, 0, (int)y.size(), len.data(), mean.data(), nullptr, 0, nullptr);
        CHECK(k == (int)seg.lengths.size());
        for (int i = 0; i < k && i < (int)seg.lengths.size(); ++i) { CHECK(len[i] == seg.lengths[i]); CHECK(mean[i] == seg.means[i]); }
        // both engines must now be at the same position
        for (int i = 0; i < 5; ++i) CHECK(rng() == orc_rng_u64(&orng));
    }
    {   // Weighted_SegmentDriver_MatchesDNAcopy_SimpleCase (tests/cbs_test.cpp:309-330), all four parameter rows
        std::vector<double> xw, ww;
        for (int i = 0; i < 15; ++i) { xw.push_back(0.0); ww.push_back(1.0); }
        for (int i = 0; i < 15; ++i) { xw.push_back(2.0); ww.push_back(0.5); }
        for (int i = 0; i < 15; ++i) { xw.push_back(-1.5); ww.push_back(2.0); }
        for (int i = 0; i < 15; ++i) { xw.push_back(0.0); ww.push_back(1.0); }
        const struct { double alpha; int nperm; bool hybrid; int mw; } cases[] = {{0.01, 200, false, 2}, {0.05, 100, false, 3},
                                                                                 {0.01, 200, true, 2}, {0.05, 100, true, 3}};
        for (const auto& tc : cases) {
            std::mt19937_64 rng(1);
            std::vector<int> sbdry((tc.nperm + 1) * (tc.nperm + 2) / 2 + 2, tc.nperm + 1);
            const auto seg = cbs_gpu::segment_weighted(xw, ww, tc.alpha, tc.nperm, tc.hybrid, tc.mw, 25, 200, 0.05, sbdry, 1e-6, rng, false, 0.05);
            CHECK(seg.lengths.size() == 4);
            if (seg.lengths.size() == 4) {
                CHECK(seg.lengths[0] == 15 && seg.lengths[1] == 15 && seg.lengths[2] == 15 && seg.lengths[3] == 15);
                CHECK(std::fabs(seg.means[0]) < 1e-9 && std::fabs(seg.means[1] - 2.0) < 1e-9 && std::fabs(seg.means[2] + 1.5) < 1e-9 &&
                      std::fabs(seg.means[3]) < 1e-9);
            }
        }
        // noisy weighted vector against the oracle, engine position included
        std::mt19937_64 g(5);
        std::normal_distribution<double> nz(0.0, 0.2);
        std::uniform_real_distribution<double> uw(0.5, 2.0);
        std::vector<double> y(700), w(700);
        for (size_t i = 0; i < y.size(); ++i) { y[i] = (double)(float)(nz(g) + ((i >= 200 && i < 330) ? 0.4 : 0.0)); w[i] = uw(g); }
        std::mt19937_64 rng(11);
        orc_rng orng;
        orc_rng_seed_mt(&orng, 11);
        const int nperm = 300;
        std::vector<int> sbdry(2000, nperm + 1);
        const auto seg = cbs_gpu::segment_weighted(y, w, 0.01, nperm, false, 2, 25, 200, 0.05, sbdry, 1e-6, rng);
        orc_seg_opts o = {0, 0.01, nperm, 0, 2, 25, 200, 0.05, 1e-6, 0, 0.05};
        std::vector<int> len(y.size());
        std::vector<double> mean(y.size());
        const int k = orc_segment_weighted(y.data(), w.data(), (int)y.size(), &o, &orng, 11, 0, (int)y.size(), len.data(), mean.data());
        CHECK(k == (int)seg.lengths.size());
        for (int i = 0; i < k && i < (int)seg.lengths.size(); ++i) { CHECK(len[i] == seg.lengths[i]); CHECK(mean[i] == seg.means[i]); }
        for (int i = 0; i < 5; ++i) CHECK(rng() == orc_rng_u64(&orng));
    }
    {   // smooth: size mismatch and negative region throw std::invalid_argument (smooth.cpp:125-126)
        bool threw = false;
        try { cbs_gpu::smooth({1.0, 2.0}, {1}); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw);
        threw = false;
        try { cbs_gpu::smooth({1.0, 2.0}, {1, 1}, -1); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw);
        std::vector<double> v(500);
        std::vector<int> lab(500, 1);
        std::mt19937_64 g(3);
        std::normal_distribution<double> nz(0.0, 0.2);
        for (auto& e : v) e = nz(g);
        v[100] += 4.0; v[300] -= 5.0; v[7] = NAN;
        const auto got = cbs_gpu::smooth(v, lab);
        std::vector<double> want(v.size());
        CHECK(orc_smooth(v.data(), lab.data(), (int64_t)v.size(), 10, 4.0, 2.0, 0.025, want.data()) == 0);
        for (size_t i = 0; i < v.size(); ++i) CHECK((got[i] == want[i]) || (std::isnan(got[i]) && std::isnan(want[i])));
        CHECK(got[100] != v[100]);
    }
    {   // cngpld::summarize_cn, tests/cngpld_test.cpp:46-83: cases 1 (amp and del) and 4 (explicit positions, empty overlap)
        struct Seg { unsigned long start, end; float value; };
        const std::vector<Seg> c1 = {{10, 25, 0.2f}, {20, 35, 0.8f}, {30, 40, 0.1f}};
        const auto amp = cngpld_gpu::summarize_cn(c1, 1, 0.5);
        const unsigned long p1[6] = {10, 20, 25, 30, 35, 40};
        const double v1[6] = {0, 1.11277046424623, 1.11277046424623, 1.11277046424623, 1.11277046424623, 0};
        CHECK(amp.size() == 6);
        for (size_t i = 0; i < amp.size() && i < 6; ++i) { CHECK(amp[i].pos == p1[i]); CHECK(std::fabs(amp[i].value - v1[i]) <= 1e-7 * v1[i]); }
        for (const auto& pt : cngpld_gpu::summarize_cn(c1, -1, 0.5)) CHECK(pt.value == 0.0);
        const std::vector<Seg> c4 = {{100, 120, 1.0f}, {200, 220, -1.0f}};
        const std::vector<unsigned long> at = {50, 100, 120, 150, 200, 220, 250};
        const auto e4 = cngpld_gpu::summarize_cn(c4, 1, 0.5, &at);
        const double v4[7] = {0, 2.71828182845905, 2.71828182845905, 0, 0, 0, 0};
        CHECK(e4.size() == 7);
        for (size_t i = 0; i < e4.size() && i < 7; ++i) { CHECK(e4[i].pos == at[i]); CHECK(std::fabs(e4[i].value - v4[i]) <= 1e-7 * v4[i]); }
        bool threw = false;
        try { cngpld_gpu::summarize_cn(c1, 2, 0.5); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw);
        const std::vector<unsigned long> none;
        CHECK(cngpld_gpu::summarize_cn(c1, 1, 0.5, &none).empty());
    }
    std::printf(fails ? "shim_test: %d failures\n" : "shim_test: ok\n", fails);
    return fails ? 1 : 0;
}
