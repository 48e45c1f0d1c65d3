// cn_reader.hpp -- reader of raw log-ratio matrices (.cn) for the `cna segment` path on the B200 host side.
//
// Restates RawSampleSet<float>::_read (/root/reference lib/RawSampleSet.hpp:217-285) with the per-chromosome position
// sort of :332-386, lib/parse.hpp:20-26 (std::from_chars, whole field must parse) and the chromosome names of
// lib/global.hpp:62-90.  read_cn() is the sequential form; read_cn_parallel() (once CBS runs on the GPU,
// parsing the text dominates the wall clock of `cna segment`) reads the file into memory, cuts it at line ends into one
// slice per thread, parses the slices concurrently into thread-local columns and concatenates them in file order, so
// its result is identical to read_cn() for every input (tests/cpp/reader_test.cpp).
#pragma once
#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

namespace cnio {

constexpr int kChromosomes = 24;

// 1..24, 0 = unknown (lib/global.hpp:62-90)
inline int chromosome_index(const std::string& name) {
    static const std::map<std::string, int> table = [] {
        std::map<std::string, int> m;
        for (int i = 1; i <= kChromosomes; ++i) {
            m[std::to_string(i)] = i;
            m["chr" + std::to_string(i)] = i;
        }
        m["X"] = 23; m["Y"] = 24; m["chrX"] = 23; m["chrY"] = 24;
        return m;
    }();
    const auto it = table.find(name);
    return it == table.end() ? 0 : it->second;
}

// tab separated fields; an empty trailing field counts (lib/parse.cpp:27-38)
struct Fields {
    std::string_view line;
    size_t pos = 0;
    explicit Fields(const std::string& s) : line(s) {}
    bool next(std::string_view& f) {
        if (pos > line.size()) return false;
        const size_t start = pos;
        while (pos < line.size() && line[pos] != '\t') ++pos;
        f = line.substr(start, pos - start);
        pos = (pos < line.size()) ? pos + 1 : line.size() + 1;
        return true;
    }
};

template <class T>
bool parse_number(std::string_view text, T& value) {  // lib/parse.hpp:20-26
    const char* b = text.data();
    const char* e = b + text.size();
    const auto r = std::from_chars(b, e, value);
    return r.ec == std::errc() && r.ptr == e;
}

struct RawMatrix {
    std::vector<std::string> sample_names;
    std::vector<unsigned long> positions[kChromosomes];
    std::vector<std::vector<float>> values[kChromosomes];  // [chrom][sample][marker]
};

// every chromosome ordered by position (RawSampleSet<V>::sort, lib/RawSampleSet.hpp:332-386).  A sample column whose length
// differs from the marker count (a row had unparsable or missing fields) is indexed out of bounds by the reference; here
// it is an error.
inline void sort_by_position(RawMatrix& m, int nthreads) {
    auto sort_chrom = [&](int c) {
        const size_t n = m.positions[c].size();
        // (position, row) pairs ordered on the position alone by std::sort, as lib/RawSampleSet.hpp:361-368 with
        // lib/global.hpp:170-173: rows that share a position come out in whatever order libstdc++'s introsort leaves
        // them (insertion order up to 16 rows, not stable beyond), and the same call gives the same order here
        std::vector<std::pair<unsigned long, size_t>> order;
        order.reserve(n);
        for (size_t i = 0; i < n; ++i) order.emplace_back(m.positions[c][i], i);
        std::sort(order.begin(), order.end(),
                  [](const std::pair<unsigned long, size_t>& a, const std::pair<unsigned long, size_t>& b) { return a.first < b.first; });
        std::vector<unsigned long> p(n);
        for (size_t i = 0; i < n; ++i) p[i] = order[i].first;
        m.positions[c].swap(p);
        for (auto& sv : m.values[c]) {
            if (sv.size() != n) throw std::runtime_error("sample column count differs from marker count (unparsable fields?)");
            std::vector<float> v(n);
            for (size_t i = 0; i < n; ++i) v[i] = sv[order[i].second];
            sv.swap(v);
        }
    };
    if (nthreads <= 1) { for (int c = 0; c < kChromosomes; ++c) sort_chrom(c); return; }
    std::vector<std::thread> th;
    std::vector<std::string> errs((size_t)kChromosomes);
    for (int c = 0; c < kChromosomes; ++c)
        th.emplace_back([&, c] { try { sort_chrom(c); } catch (const std::exception& e) { errs[(size_t)c] = e.what(); } });
    for (auto& t : th) t.join();
    for (const auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
}

inline RawMatrix read_cn(const std::string& path) {
    std::ifstream file(path);
    if (!file.is_open()) throw std::runtime_error("Failed to open input file '" + path + "'.");
    RawMatrix m;
    std::string line;
    size_t line_no = 0;
    for (;;) {
        std::getline(file, line);
        if (file.eof()) break;  // as the reference: a last line without newline is not processed
        ++line_no;
        Fields fields(line);
        std::string_view f;
        if (line_no == 1) {
            for (int i = 0; i < 3 && fields.next(f); ++i) {}
            while (fields.next(f)) m.sample_names.emplace_back(f);
            for (auto& v : m.values) v.assign(m.sample_names.size(), {});
            continue;
        }
        std::string chrom_name;
        unsigned long pos = 0;
        if (!fields.next(f)) continue;  // marker name
        if (!fields.next(f)) continue;
        chrom_name.assign(f);
        if (!fields.next(f) || !parse_number(f, pos)) continue;
        const int chr = chromosome_index(chrom_name);
        if (chr == 0) continue;  // unknown chromosome: row ignored
        m.positions[chr - 1].push_back(pos);
        size_t s = 0;
        while (fields.next(f)) {
            float v;
            if (!parse_number(f, v)) continue;  // unparsable fields are skipped, later columns shift
            if (s < m.sample_names.size()) m.values[chr - 1][s].push_back(v);
            ++s;
        }
    }
    sort_by_position(m, 1);
    return m;
}


// one data line (without its '\n') into the columns of `m`; same rules as the loop body of read_cn
inline void parse_cn_line(std::string_view line, size_t n_samples, RawMatrix& m) {
    size_t pos0 = 0;
    auto next = [&](std::string_view& f) {
        if (pos0 > line.size()) return false;
        const size_t start = pos0;
        while (pos0 < line.size() && line[pos0] != '\t') ++pos0;
        f = line.substr(start, pos0 - start);
        pos0 = (pos0 < line.size()) ? pos0 + 1 : line.size() + 1;
        return true;
    };
    std::string_view f;
    unsigned long pos = 0;
    if (!next(f)) return;  // marker name
    if (!next(f)) return;
    const std::string chrom_name(f);
    if (!next(f) || !parse_number(f, pos)) return;
    const int chr = chromosome_index(chrom_name);
    if (chr == 0) return;
    m.positions[chr - 1].push_back(pos);
    size_t s = 0;
    while (next(f)) {
        float v;
        if (!parse_number(f, v)) continue;
        if (s < n_samples) m.values[chr - 1][s].push_back(v);
        ++s;
    }
}

inline RawMatrix read_cn_parallel(const std::string& path, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("Failed to open input file '" + path + "'.");
    std::fseek(fp, 0, SEEK_END);
    const long size = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::string buf((size_t)std::max(0L, size), '\0');
    const size_t got = size > 0 ? std::fread(buf.data(), 1, (size_t)size, fp) : 0;
    std::fclose(fp);
    buf.resize(got);
    RawMatrix m;
    // only lines that end with '\n' are processed (the reference stops at eof inside getline)
    const size_t last_nl = buf.rfind('\n');
    if (last_nl == std::string::npos) return m;
    const size_t end = last_nl + 1;
    const size_t hdr_end = buf.find('\n');
    {
        // (Fields keeps a view of its argument: parse the header from a local copy)
        const std::string header = buf.substr(0, hdr_end);
        Fields hf(header);
        std::string_view f;
        for (int i = 0; i < 3 && hf.next(f); ++i) {}
        while (hf.next(f)) m.sample_names.emplace_back(f);
        for (auto& v : m.values) v.assign(m.sample_names.size(), {});
    }
    const size_t body = hdr_end + 1;
    const size_t n_samples = m.sample_names.size();
    // slice boundaries: byte offsets moved forward to the next line start
    std::vector<size_t> cut((size_t)nthreads + 1, end);
    cut[0] = body;
    for (int t = 1; t < nthreads; ++t) {
        size_t at = body + (end - body) / (size_t)nthreads * (size_t)t;
        if (at < body) at = body;
        const size_t nl = buf.find('\n', at);
        cut[(size_t)t] = (nl == std::string::npos || nl + 1 > end) ? end : nl + 1;
        if (cut[(size_t)t] < cut[(size_t)t - 1]) cut[(size_t)t] = cut[(size_t)t - 1];
    }
    std::vector<RawMatrix> part((size_t)nthreads);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t] {
            RawMatrix& pm = part[(size_t)t];
            for (auto& v : pm.values) v.assign(n_samples, {});
            size_t at = cut[(size_t)t];
            const size_t stop = cut[(size_t)t + 1];
            while (at < stop) {
                const char* nl = (const char*)std::memchr(buf.data() + at, '\n', stop - at);
                const size_t e = nl ? (size_t)(nl - buf.data()) : stop;
                parse_cn_line(std::string_view(buf.data() + at, e - at), n_samples, pm);
                at = e + 1;
            }
        });
    for (auto& t : th) t.join();
    // concatenate the slices in file order (threads over chromosomes)
    auto merge_chrom = [&](int c) {
        size_t n = 0;
        for (const auto& pm : part) n += pm.positions[c].size();
        m.positions[c].reserve(n);
        for (const auto& pm : part) m.positions[c].insert(m.positions[c].end(), pm.positions[c].begin(), pm.positions[c].end());
        for (size_t s = 0; s < n_samples; ++s) {
            auto& dst = m.values[c][s];
            size_t ns = 0;
            for (const auto& pm : part) ns += pm.values[c][s].size();
            dst.reserve(ns);
            for (const auto& pm : part) dst.insert(dst.end(), pm.values[c][s].begin(), pm.values[c][s].end());
        }
    };
    {
        std::vector<std::thread> mt;
        for (int c = 0; c < kChromosomes; ++c) mt.emplace_back([&, c] { merge_chrom(c); });
        for (auto& t : mt) t.join();
    }
    part.clear();
    sort_by_position(m, nthreads);
    return m;
}

}  // namespace cnio
