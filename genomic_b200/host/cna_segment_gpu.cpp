// cna_segment_gpu -- the `cna segment` command on the B200 path.
//
// C++ host driver equivalent to Segment::run / Segment::segment_raw (src/cna_segment.hpp:45-62,
// :127-159): reads a raw log-ratio matrix (.cn), checks log scale, hands every (sample, chromosome)
// unit to libcbs_cuda.so through the C ABI (one batched call: smoothing + CBS on the device, ONE
// std::mt19937_64(1) stream shared serially by all units exactly as the reference does), and writes
// the segment table (.seg) with the reference's row layout.
//
//   reader : lib/RawSampleSet.hpp:217-285 (+ per-chromosome position sort :332-386), lib/parse.hpp:20-26,
//            chromosome names lib/global.hpp:62-90
//   writer : lib/SegmentedSampleSet.hpp:519-535 (float state, default ostream precision)
//   options: src/cna_segment.hpp:23-42, defaults :67-79
//
// usage: cna_segment_gpu [options] <raw sample matrix file> [<output segmentation file>]
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "cbs_gpu.h"

#include "cn_reader.hpp"

namespace {

using namespace cnio;

// src/cna_segment.hpp:109-125
void ensure_log_scale(const RawMatrix& m) {
    bool neg = false, pos = false;
    for (const auto& chrom : m.values)
        for (const auto& sample : chrom)
            for (float v : sample) {
                if (!std::isfinite(v)) continue;
                if (v < 0) neg = true;
                if (v > 0) pos = true;
            }
    if (!(neg && pos)) throw std::invalid_argument("Input does not appear to be in log scale: expected both negative and positive values.");
}

std::string filestem(const std::string& s) {  // lib/global.cpp name::filestem
    size_t start = s.find_last_of('/');
    start = (start == std::string::npos) ? 0 : start + 1;
    const size_t end = s.find_last_of('.');
    return s.substr(start, (end == std::string::npos) ? std::string::npos : end - start);
}

}  // namespace

int main(int argc, char** argv) {
    try {
        cbs_gpu_params p;
        cbs_gpu_default_params(&p);
        std::string input, output;
        int device = 0, chain_opt = -1;
        bool timing = false;
        int io_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));  // parser threads (cn_reader.hpp)
        std::vector<std::string> positional;
        auto value_of = [&](int& i, const std::string& arg, const std::string& name, std::string& out) -> bool {
            if (arg == name) { if (i + 1 >= argc) throw std::invalid_argument("missing value for " + name); out = argv[++i]; return true; }
            if (arg.rfind(name + "=", 0) == 0) { out = arg.substr(name.size() + 1); return true; }
            return false;
        };
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            std::string v;
            if (a == "--help") {
                std::cout << "usage:  cna_segment_gpu [options] <raw sample matrix file> <output segmentation file>\n"
                             "  --alpha --nperm --min_width --kmax --nmin --eta --trim --smooth_region --outlier_sd_scale\n"
                             "  --smooth_sd_scale --hybrid --undo_prune --undo_prune_cutoff  (as `cna segment`)\n"
                             "  --device N   --rng mt|philox   --seed S   --io_threads T (parser threads, default: all cores up to 16)\n"
                             "  --chain 0|1  MT replay: 1 = ONE engine shared serially by all units, as `cna segment` does (default);\n"
                             "               0 = a fresh std::mt19937_64(seed) per (sample, chromosome): units are independent and run together\n"
                             "  --timing     phase times as one JSON line on stderr   --format cn (accepted for compatibility)\n";
                return 0;
            } else if (value_of(i, a, "--input", v) || value_of(i, a, "-i", v)) input = v;
            else if (value_of(i, a, "--output", v) || value_of(i, a, "-o", v)) output = v;
            else if (value_of(i, a, "--alpha", v)) p.alpha = std::stod(v);
            else if (value_of(i, a, "--nperm", v)) p.nperm = std::stoi(v);
            else if (value_of(i, a, "--min_width", v)) p.min_width = std::stoi(v);
            else if (value_of(i, a, "--kmax", v)) p.kmax = std::stoi(v);
            else if (value_of(i, a, "--nmin", v)) p.nmin = std::stoi(v);
            else if (value_of(i, a, "--eta", v)) p.eta = std::stod(v);
            else if (value_of(i, a, "--trim", v)) p.trim = std::stod(v);
            else if (value_of(i, a, "--smooth_region", v)) p.smooth_region = std::stoi(v);
            else if (value_of(i, a, "--outlier_sd_scale", v)) p.outlier_sd_scale = std::stod(v);
            else if (value_of(i, a, "--smooth_sd_scale", v)) p.smooth_sd_scale = std::stod(v);
            else if (value_of(i, a, "--hybrid", v)) p.hybrid = (v == "1" || v == "true");
            else if (value_of(i, a, "--undo_prune", v)) p.undo_prune = (v == "1" || v == "true");
            else if (value_of(i, a, "--undo_prune_cutoff", v)) p.undo_prune_cutoff = std::stod(v);
            else if (value_of(i, a, "--device", v)) device = std::stoi(v);
            else if (value_of(i, a, "--io_threads", v)) io_threads = std::max(1, std::stoi(v));
            else if (value_of(i, a, "--seed", v)) p.seed = std::stoull(v);
            else if (value_of(i, a, "--format", v)) { if (v != "cn") throw std::invalid_argument("segment command currently supports raw log-ratio matrices only."); }
            else if (value_of(i, a, "--chain", v)) chain_opt = (v == "1" || v == "true") ? 1 : 0;
            else if (a == "--timing") timing = true;
            else if (value_of(i, a, "--rng", v)) { p.rng_mode = (v == "philox") ? CBS_GPU_RNG_PHILOX : CBS_GPU_RNG_MT19937_64; p.chain = (v == "philox") ? 0 : 1; }
            else if (!a.empty() && a[0] == '-') throw std::invalid_argument("unknown option " + a);
            else positional.push_back(a);
        }
        if (chain_opt >= 0 && p.rng_mode == CBS_GPU_RNG_MT19937_64) p.chain = chain_opt;
        if (input.empty() && !positional.empty()) { input = positional.front(); positional.erase(positional.begin()); }
        if (output.empty() && !positional.empty()) output = positional.front();
        if (input.empty()) throw std::invalid_argument("Input file not specified.");
        {
            const size_t dot = input.find_last_of('.');
            const std::string ext = (dot == std::string::npos || dot == input.size() - 1) ? input : input.substr(dot + 1);
            if (ext != "cn") throw std::invalid_argument("segment command currently supports raw log-ratio matrices only.");
        }
        if (output.empty()) output = filestem(input) + ".seg";

        const auto t_start = std::chrono::steady_clock::now();
        // the CUDA context comes up (driver initialisation, module load) on a second thread while the text is parsed
        cbs_gpu_ctx* ctx = nullptr;
        int create_rc = CBS_GPU_OK;
        double create_ms = 0.0;
        std::thread creator([&] {
            const int ids[1] = {device};
            const auto c0 = std::chrono::steady_clock::now();
            create_rc = cbs_gpu_create(ids, 1, &ctx);
            create_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count();
        });
        struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{creator};
        const RawMatrix m = io_threads > 1 ? read_cn_parallel(input, io_threads) : read_cn(input);
        ensure_log_scale(m);
        const auto t_read = std::chrono::steady_clock::now();

        // units in the reference's order: samples in file order, chromosomes 1..24, empty ones skipped
        std::vector<float> values;
        std::vector<int64_t> off{0};
        struct Unit { size_t sample; int chrom; };
        std::vector<Unit> units;
        for (size_t s = 0; s < m.sample_names.size(); ++s)
            for (int c = 0; c < kChromosomes; ++c) {
                const auto& v = m.values[c][s];
                if (v.empty()) continue;
                values.insert(values.end(), v.begin(), v.end());
                off.push_back((int64_t)values.size());
                units.push_back({s, c});
            }

        const auto t_pack = std::chrono::steady_clock::now();
        creator.join();
        const auto t_ctx = std::chrono::steady_clock::now();
        if (create_rc != CBS_GPU_OK) throw std::runtime_error("no usable CUDA device (there is no CPU fallback)");
        cbs_gpu_result* res = nullptr;
        const int rc = cbs_gpu_segment_batch(ctx, values.data(), CBS_GPU_F32, CBS_GPU_HOST, off.data(), nullptr, (int32_t)units.size(), &p, &res);
        if (rc != CBS_GPU_OK) {
            const std::string msg = cbs_gpu_last_error(ctx);
            cbs_gpu_destroy(ctx);
            if (rc == CBS_GPU_ERR_INVALID) throw std::invalid_argument(msg);
            throw std::runtime_error(msg);
        }

        const auto t_gpu = std::chrono::steady_clock::now();
        std::ofstream out(output);
        if (!out.is_open()) throw std::runtime_error("Failed to open output file '" + output + "'.");
        out << "sample\tchromosome\tstart\tend\tcount\tstate" << '\n';
        for (size_t u = 0; u < units.size(); ++u) {
            const auto& pos = m.positions[units[u].chrom];
            size_t start = 0;
            for (int64_t k = res->seg_offsets[u]; k < res->seg_offsets[u + 1]; ++k) {
                const size_t len = (size_t)res->lengths[k];
                if (len == 0) continue;
                const size_t end = start + len - 1;
                const float state = (float)res->means[k];  // Segment<rvalue>: rvalue = float (lib/typedefs.h:20)
                out << m.sample_names[units[u].sample] << '\t' << (units[u].chrom + 1) << '\t' << pos[start] << '\t' << pos[end]
                    << '\t' << (unsigned long)len << '\t' << state << '\n';
                start += len;
            }
        }
        out.flush();
        const auto t_write = std::chrono::steady_clock::now();
        if (timing) {
            auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            long long markers = 0;
            for (size_t u = 0; u < units.size(); ++u) markers += off[u + 1] - off[u];
            std::fprintf(stderr,
                         "{\"samples\": %zu, \"units\": %zu, \"markers_x_samples\": %lld, \"segments\": %lld, \"io_threads\": %d, "
                         "\"read_parse_sort_ms\": %.1f, \"pack_ms\": %.1f, \"gpu_create_ms\": %.1f, \"gpu_create_wait_ms\": %.1f, \"gpu_call_ms\": %.1f, \"gpu_h2d_ms\": %.1f, \"gpu_smooth_ms\": %.1f, "
                         "\"gpu_segment_ms\": %.1f, \"gpu_d2h_ms\": %.1f, \"write_seg_ms\": %.1f, \"total_ms\": %.1f}\n",
                         m.sample_names.size(), units.size(), markers, (long long)res->n_segments, io_threads, ms(t_start, t_read),
                         ms(t_read, t_pack), create_ms, ms(t_pack, t_ctx), ms(t_ctx, t_gpu), res->ms_h2d, res->ms_smooth, res->ms_segment, res->ms_d2h, ms(t_gpu, t_write),
                         ms(t_start, t_write));
        }
        cbs_gpu_result_free(res);
        cbs_gpu_destroy(ctx);
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;  // src/cna.cpp:96-99
        return 1;
    }
}
