// CPU test and timing of the .cn readers (genomic_b200/host/cn_reader.hpp): the parallel reader must return exactly what
// the sequential restatement of RawSampleSet<float>::_read returns.
//   reader_test <tmp.cn> <markers per chromosome> <samples> <threads>   -> one JSON line
#include <chrono>
#include <cstdio>
#include <random>

#include "../../genomic_b200/host/cn_reader.hpp"

static bool same(const cnio::RawMatrix& a, const cnio::RawMatrix& b) {
    if (a.sample_names != b.sample_names) return false;
    for (int c = 0; c < cnio::kChromosomes; ++c) {
        if (a.positions[c] != b.positions[c]) return false;
        if (a.values[c].size() != b.values[c].size()) return false;
        for (size_t s = 0; s < a.values[c].size(); ++s) {
            const auto& x = a.values[c][s];
            const auto& y = b.values[c][s];
            if (x.size() != y.size()) return false;
            for (size_t i = 0; i < x.size(); ++i)
                if (std::memcmp(&x[i], &y[i], sizeof(float)) != 0) return false;  // bit pattern (NaN included)
        }
    }
    return true;
}

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: reader_test file markers samples threads\n"); return 2; }
    const std::string path = argv[1];
    const int per_chrom = std::atoi(argv[2]), samples = std::atoi(argv[3]), threads = std::atoi(argv[4]);
    {
        std::FILE* f = std::fopen(path.c_str(), "w");
        if (!f) return 2;
        std::fprintf(f, "marker\tchromosome\tposition");
        for (int s = 0; s < samples; ++s) std::fprintf(f, "\tS%d", s);
        std::fprintf(f, "\n");
        std::mt19937_64 g(7);
        std::normal_distribution<float> nz(0.f, 0.2f);
        long id = 0;
        for (int c = 1; c <= 25; ++c) {  // chromosome "25" is unknown: its rows are ignored
            for (int i = 0; i < per_chrom; ++i) {
                const char* pre = (c % 2) ? "chr" : "";
                // positions deliberately not monotone: the reader sorts them
                const unsigned long pos = 1000ul + (unsigned long)((i * 7919L) % per_chrom) * 10ul + (unsigned long)(i % 3 == 0);
                if (c == 23) std::fprintf(f, "m%ld\t%sX\t%lu", id++, pre, pos);
                else std::fprintf(f, "m%ld\t%s%d\t%lu", id++, pre, c, pos);
                for (int s = 0; s < samples; ++s) {
                    const float v = nz(g);
                    if (i == 5 && s == 1) std::fprintf(f, "\tnan");
                    else if (i == 6 && s == 0) std::fprintf(f, "\t1e-3");
                    else std::fprintf(f, "\t%.6g", v);
                }
                std::fprintf(f, "\n");
            }
        }
        std::fprintf(f, "tail\t1\t5\t0.5");  // no newline: not processed
        std::fclose(f);
    }
    using clk = std::chrono::steady_clock;
    double s_seq = 1e30, s_par = 1e30;
    bool ok = true;
    cnio::RawMatrix a;
    for (int rep = 0; rep < 3; ++rep) {  // best of 3: shared hosts are noisy
        const auto t0 = clk::now();
        a = cnio::read_cn(path);
        const auto t1 = clk::now();
        const cnio::RawMatrix b = cnio::read_cn_parallel(path, threads);
        const auto t2 = clk::now();
        ok = ok && same(a, b);
        s_seq = std::min(s_seq, std::chrono::duration<double>(t1 - t0).count());
        s_par = std::min(s_par, std::chrono::duration<double>(t2 - t1).count());
    }
    ok = ok && same(a, cnio::read_cn_parallel(path, 1)) && same(a, cnio::read_cn_parallel(path, 3));
    size_t values = 0;
    for (int c = 0; c < cnio::kChromosomes; ++c) for (const auto& v : a.values[c]) values += v.size();
    std::printf("{\"identical\": %s, \"values\": %zu, \"threads\": %d, \"sequential_s\": %.3f, \"parallel_s\": %.3f, "
                "\"sequential_values_per_s\": %.0f, \"parallel_values_per_s\": %.0f}\n",
                ok ? "true" : "false", values, threads, s_seq, s_par, values / s_seq, values / s_par);
    std::remove(path.c_str());
    return ok ? 0 : 1;
}
