// synth.cpp -- synthetic SNP6-scale inputs, shared by tests, bench.py and
// the CPU baseline so that every arm segments the SAME buffers.  libstdc++ distributions are
// used on purpose (they are not portable across standard libraries): generate once, here.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

namespace {
// markers per chromosome, proportional to hg19 lengths, sum = 1,800,000; chrY empty
const int kChromSizes[23] = {147726, 144166, 117469, 113316, 107383, 101450, 94331, 86618, 83652, 80686, 80092, 79499,
                             68227,  63481,  61107,  53395,  48055,  46276,  35003, 37376, 28477, 30257, 91958};
const double kLevels[6] = {-1.0, -0.4, 0.0, 0.3, 0.58, 1.0};
}  // namespace

extern "C" {

int synth_n_chrom(void) { return 23; }
int synth_chrom_size(int chrom1) { return (chrom1 >= 1 && chrom1 <= 23) ? kChromSizes[chrom1 - 1] : 0; }

// One (sample, chromosome) unit of n markers: piecewise-constant log2 ratio + N(0, 0.2^2) noise,
// stored as float32.  outliers != 0 injects one +-(3+|N(0,1)|) spike every 1000 markers.
// n need not be the SNP6 size (tests use small n); segments are at least min(50, n/8) markers apart.
void synth_unit(uint64_t sample, int chrom1, int n, int outliers, float* out) {
    std::mt19937_64 g(20260101ULL + 1000ULL * sample + (uint64_t)chrom1);
    std::poisson_distribution<int> pois(3.0);
    int nseg = 1 + std::min(7, pois(g));
    const int spacing = std::max(1, std::min(50, n / 8));
    std::vector<int> cuts;
    if (n > 2 * spacing) {
        std::uniform_int_distribution<int> pos(spacing, n - spacing);
        int tries = 0;
        while ((int)cuts.size() < nseg - 1 && tries < 1000) {
            ++tries;
            const int c = pos(g);
            bool ok = true;
            for (int e : cuts) if (std::abs(e - c) < spacing) ok = false;
            if (ok) cuts.push_back(c);
        }
    }
    std::sort(cuts.begin(), cuts.end());
    cuts.push_back(n);
    std::uniform_int_distribution<int> lev(0, 5);
    std::normal_distribution<double> noise(0.0, 0.2);
    int prev_level = -1, start = 0;
    for (int end : cuts) {
        int l = lev(g);
        while (l == prev_level) l = lev(g);
        prev_level = l;
        for (int i = start; i < end; ++i) out[i] = (float)(kLevels[l] + noise(g));
        start = end;
    }
    if (outliers) {
        std::normal_distribution<double> z(0.0, 1.0);
        const int phase = (int)(g() % 1000ULL);
        double sign = 1.0;
        for (int i = phase; i < n; i += 1000) {
            out[i] = (float)((double)out[i] + sign * (3.0 + std::fabs(z(g))));
            sign = -sign;
        }
    }
}

// config 5 units: pure null N(0, 0.2^2), optionally with `shift` added from marker shift_at on
void synth_null_unit(uint64_t seed, int n, int shift_at, double shift, float* out) {
    std::mt19937_64 g(seed);
    std::normal_distribution<double> noise(0.0, 0.2);
    for (int i = 0; i < n; ++i) out[i] = (float)(noise(g) + ((shift_at >= 0 && i >= shift_at) ? shift : 0.0));
}

}  // extern "C"
