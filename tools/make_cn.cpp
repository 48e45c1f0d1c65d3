// make_cn.cpp -- writes a synthetic SNP6-scale cohort (genomic_b200/host/synth.cpp) as a `.cn` raw matrix
// (lib/RawSampleSet.hpp:217-263 layout: marker, chromosome, position, one column per sample) for end-to-end runs of
// cna_segment_gpu.   g++ -O2 -std=c++17 -o tools/make_cn tools/make_cn.cpp -Lgenomic_b200 -l:libsynth.so -Wl,-rpath,'$ORIGIN/../genomic_b200'
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" int synth_chrom_size(int chrom1);
extern "C" void synth_unit(unsigned long long sample, int chrom1, int n, int outliers, float* out);
int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: make_cn <out.cn> <samples> [scale]\n"); return 2; }
    const int S = std::atoi(argv[2]);
    const double scale = argc > 3 ? std::atof(argv[3]) : 1.0;
    FILE* f = std::fopen(argv[1], "w");
    if (!f) return 1;
    static char buf[1 << 20];
    std::setvbuf(f, buf, _IOFBF, sizeof(buf));
    std::fprintf(f, "marker\tchromosome\tposition");
    for (int s = 0; s < S; ++s) std::fprintf(f, "\tS%04d", s);
    std::fprintf(f, "\n");
    for (int c = 1; c <= 23; ++c) {
        int n = (int)(synth_chrom_size(c) * scale + 0.5);
        if (n < 8) n = 8;
        std::vector<std::vector<float>> col((size_t)S, std::vector<float>((size_t)n));
        for (int s = 0; s < S; ++s) synth_unit((unsigned long long)s, c, n, 0, col[(size_t)s].data());
        for (int k = 0; k < n; ++k) {
            std::fprintf(f, "m%d_%d\t%d\t%d", c, k, c, 1000 * (k + 1));
            for (int s = 0; s < S; ++s) std::fprintf(f, "\t%.7g", (double)col[(size_t)s][(size_t)k]);
            std::fputc('\n', f);
        }
    }
    std::fclose(f);
    return 0;
}
