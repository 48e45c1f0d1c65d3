// Adversarial inputs for the .cn reader (genomic_b200/host/cn_reader.hpp) against the rules of the reference's reader,
// derived by hand from /root/reference lib/RawSampleSet.hpp:217-285 (_read, readSampleValues), :332-386 (sort),
// lib/parse.hpp:20-26 (std::from_chars, whole field) and lib/global.hpp:62-90 (chromosome names).  The reference reader
// itself needs boost and a generated config.h, so it is not compiled here; every expectation below cites its rule.
//   reader_rules_test <tmp dir>   -> prints the failed checks, exit code 0 when none
#include <cmath>
#include <cstdio>
#include <functional>

#include "../../genomic_b200/host/cn_reader.hpp"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

static std::string write_file(const std::string& dir, const char* name, const std::string& text) {
    const std::string path = dir + "/" + name;
    std::FILE* f = std::fopen(path.c_str(), "wb");
    std::fwrite(text.data(), 1, text.size(), f);
    std::fclose(f);
    return path;
}

// both readers, every thread count, must agree; returns the sequential result
static cnio::RawMatrix read_all(const std::string& path) {
    const cnio::RawMatrix a = cnio::read_cn(path);
    for (int t : {1, 2, 3, 7}) {
        const cnio::RawMatrix b = cnio::read_cn_parallel(path, t);
        CHECK(a.sample_names == b.sample_names);
        for (int c = 0; c < cnio::kChromosomes; ++c) {
            CHECK(a.positions[c] == b.positions[c]);
            CHECK(a.values[c].size() == b.values[c].size());
            for (size_t s = 0; s < a.values[c].size() && s < b.values[c].size(); ++s) {
                CHECK(a.values[c][s].size() == b.values[c][s].size());
                if (a.values[c][s].size() == b.values[c][s].size() && !a.values[c][s].empty())
                    CHECK(std::memcmp(a.values[c][s].data(), b.values[c][s].data(), a.values[c][s].size() * sizeof(float)) == 0);
            }
        }
    }
    return a;
}

static bool throws(const std::function<void()>& f) {
    try { f(); } catch (const std::runtime_error&) { return true; }
    return false;
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    const std::string H = "marker\tchromosome\tposition\tA\tB\n";

    {   // rows dropped as a whole: unparsable position (_read: `!parseNumber(field, pos)) continue`), unknown chromosome
        // (`chr == 0`), fewer than three fields, empty line; names with and without "chr", X and Y (global.hpp:62-90)
        const auto m = read_all(write_file(dir, "rows.cn", H +
            "m1\t1\t100\t0.1\t0.2\n"
            "m2\t1\t12x\t9\t9\n"        // position does not parse as a whole
            "m3\t1\t-5\t9\t9\n"         // from_chars into an unsigned type rejects the sign
            "m4\t1\t 7\t9\t9\n"         // leading blank
            "m5\tMT\t50\t9\t9\n"        // unknown chromosome
            "m6\tchr25\t50\t9\t9\n"
            "\n"
            "m7\t1\n"
            "m8\tchrX\t5\t0.3\t0.4\n"
            "m9\tY\t6\t0.5\t0.6\n"
            "m10\tchr1\t50\t0.7\t0.8\n"
            "m11\t1\t99999999999\t0.9\t1.0\n"));  // positions are unsigned long: beyond 32 bits is fine
        CHECK((m.sample_names == std::vector<std::string>{"A", "B"}));
        CHECK((m.positions[0] == std::vector<unsigned long>{50, 100, 99999999999ul}));
        CHECK((m.values[0][0] == std::vector<float>{0.7f, 0.1f, 0.9f}));
        CHECK((m.values[0][1] == std::vector<float>{0.8f, 0.2f, 1.0f}));
        CHECK((m.positions[22] == std::vector<unsigned long>{5}) && (m.values[22][1] == std::vector<float>{0.4f}));
        CHECK((m.positions[23] == std::vector<unsigned long>{6}) && (m.values[23][0] == std::vector<float>{0.5f}));
        for (int c = 1; c < 22; ++c) CHECK(m.positions[c].empty());
    }
    {   // what std::from_chars<float> takes: nan, inf, -inf, exponents, hex-less plain decimals; kept as values
        const auto m = read_all(write_file(dir, "special.cn", H +
            "m1\t2\t1\tnan\tinf\n"
            "m2\t2\t2\t-inf\t1e-3\n"
            "m3\t2\t3\t-0\t.5\n"
            "m4\t2\t4\t5.\t1E2\n"));
        CHECK(m.values[1][0].size() == 4 && m.values[1][1].size() == 4);
        CHECK(std::isnan(m.values[1][0][0]) && std::isinf(m.values[1][1][0]) && m.values[1][1][0] > 0);
        CHECK(std::isinf(m.values[1][0][1]) && m.values[1][0][1] < 0 && m.values[1][1][1] == 1e-3f);
        CHECK(m.values[1][0][2] == 0.0f && std::signbit(m.values[1][0][2]) && m.values[1][1][2] == 0.5f);
        CHECK(m.values[1][0][3] == 5.0f && m.values[1][1][3] == 100.0f);
    }
    {   // fields readSampleValues skips (`if (!parseNumber(field, value)) continue;` -- the sample index does not advance,
        // so later fields shift one sample to the left and the last sample is left short): NA, empty, "+1", trailing
        // blank, out of range, "\r" of a CRLF file.  The reference then indexes the short column out of bounds in sort();
        // here the read fails with an error instead.
        for (const char* bad : {"NA", "", "+1", "0.5 ", "1e400", "0.5\r", "0x10", "1,5"}) {
            const std::string path = write_file(dir, "shift.cn", H + "m1\t3\t1\t0.1\t0.2\n" + "m2\t3\t2\t" + bad + "\t0.4\n");
            CHECK(throws([&] { cnio::read_cn(path); }));
            CHECK(throws([&] { cnio::read_cn_parallel(path, 2); }));
        }
        // the shift itself, visible when another unparsable row restores the balance is impossible (columns only shrink):
        // check it on the parsed columns before the sort with the line parser
        cnio::RawMatrix pm;
        for (auto& v : pm.values) v.assign(2, {});
        cnio::parse_cn_line("m2\t3\t2\tNA\t0.4", 2, pm);
        CHECK((pm.values[2][0] == std::vector<float>{0.4f}) && pm.values[2][1].empty());
    }
    {   // more value fields than samples: the reference writes past its sample vector; here the extra fields are dropped
        const auto m = read_all(write_file(dir, "extra.cn", H + "m1\t4\t1\t0.1\t0.2\t0.3\n"));
        CHECK((m.values[3][0] == std::vector<float>{0.1f}) && (m.values[3][1] == std::vector<float>{0.2f}));
    }
    {   // last line without '\n' is not processed (`getline; if (file.eof()) break;`); header only; empty file
        const auto m = read_all(write_file(dir, "tail.cn", H + "m1\t5\t1\t0.1\t0.2\nm2\t5\t2\t0.3\t0.4"));
        CHECK((m.positions[4] == std::vector<unsigned long>{1}));
        const auto h = read_all(write_file(dir, "header.cn", H));
        CHECK(h.sample_names.size() == 2 && h.positions[0].empty());
        const auto h2 = read_all(write_file(dir, "short_header.cn", "marker\tchromosome\n"));
        CHECK(h2.sample_names.empty());
        const auto e = read_all(write_file(dir, "empty.cn", ""));
        CHECK(e.sample_names.empty());
    }
    {   // rows sharing a position: order = std::sort on (position, row) pairs compared on the position only
        // (RawSampleSet.hpp:361-368).  Up to 16 rows libstdc++ sorts by insertion, which keeps file order; beyond that
        // the order is introsort's, reproduced here by making the same call on the same pairs.
        std::string small = H, big = H;
        for (int i = 0; i < 12; ++i) small += "s\t6\t" + std::to_string(i % 3) + "\t" + std::to_string(i) + "\t0\n";
        const auto ms = read_all(write_file(dir, "dup_small.cn", small));
        CHECK((ms.values[5][0] == std::vector<float>{0, 3, 6, 9, 1, 4, 7, 10, 2, 5, 8, 11}));
        const int n = 1000;
        std::vector<std::pair<unsigned long, size_t>> order;
        for (int i = 0; i < n; ++i) {
            const unsigned long pos = (unsigned long)((i * 37) % 11);
            big += "b\t7\t" + std::to_string(pos) + "\t" + std::to_string(i) + "\t0\n";
            order.emplace_back(pos, (size_t)i);
        }
        std::sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        const auto mb = read_all(write_file(dir, "dup_big.cn", big));
        bool same = mb.values[6][0].size() == (size_t)n, stable = true;
        for (int i = 0; same && i < n; ++i) {
            same = mb.values[6][0][(size_t)i] == (float)order[(size_t)i].second && mb.positions[6][(size_t)i] == order[(size_t)i].first;
            if (i && order[(size_t)i].first == order[(size_t)i - 1].first && order[(size_t)i].second < order[(size_t)i - 1].second) stable = false;
        }
        CHECK(same);
        std::printf("{\"duplicate_positions_order_is_stable_at_1000_rows\": %s}\n", stable ? "true" : "false");
    }
    std::printf("{\"failures\": %d}\n", failures);
    return failures ? 1 : 0;
}
