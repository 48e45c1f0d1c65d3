// mt_jump.h -- host-side mathematics for generating ONE std::mt19937_64 stream in parallel.
//
// The raw (untempered) word stream W[t] of MT19937-64 is linear over GF(2): with phi the
// characteristic polynomial (degree 19937) of its state transition,
//         sum_k phi_k W[t+k] = 0                        for every t >= 1,
// hence for g_d(x) = x^d mod phi(x):   W[t+d] = sum_i g_d[i] W[t+i].
// So 312 consecutive words at distance d are an XOR-combination of shifted windows of the 20249
// words that follow position t -- no sequential stepping.  The device generator (k_gen_lead /
// k_gen_par in kernels.cuh) produces a lead-in of LEAD words sequentially, lets CTA c compute the
// first 312 words of segment c with the polynomial g_{c*S}, and then every CTA extends its own
// segment with the ordinary recurrence.  This header computes phi (Berlekamp-Massey on one output
// bit), the table g_{c*S}, c = 1..C-1, and a self test against sequential generation.
#pragma once
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace cbsg {
namespace mtjump {

constexpr int DEG = 19937;
constexpr int PW = 312;             // words per polynomial (19968 bits >= DEG)
constexpr int SEG_LOG2 = 18;
constexpr long long SEG = 1LL << SEG_LOG2;  // words per segment
constexpr int NSEG = 512;           // segments per parallel span
constexpr int LEAD = 20480;         // sequential lead-in words (>= DEG + PW)

using Poly = std::vector<uint64_t>;  // little endian bit array, PW words (2*PW for products)

inline uint64_t twist(uint64_t a, uint64_t b, uint64_t m) {
    const uint64_t y = (a & 0xFFFFFFFF80000000ULL) | (b & 0x7FFFFFFFULL);
    return m ^ (y >> 1) ^ ((y & 1ULL) ? 0xB5026F5AA96619E9ULL : 0ULL);
}

// raw words W[0..count): W[0..312) = the seeded state, then the recurrence
inline std::vector<uint64_t> raw_stream(uint64_t seed, size_t count) {
    std::vector<uint64_t> w(count < 312 ? 312 : count);
    w[0] = seed;
    for (int k = 1; k < 312; ++k) w[k] = 6364136223846793005ULL * (w[k - 1] ^ (w[k - 1] >> 62)) + (uint64_t)k;
    for (size_t t = 312; t < w.size(); ++t) w[t] = twist(w[t - 312], w[t - 311], w[t - 156]);
    return w;
}

inline int get_bit(const Poly& p, int i) { return (int)((p[(size_t)i >> 6] >> (i & 63)) & 1ULL); }
inline void flip_bit(Poly& p, int i) { p[(size_t)i >> 6] ^= 1ULL << (i & 63); }

// Berlekamp-Massey over GF(2): minimal connection polynomial of bit sequence s (length n).
// Returns L and C (C[0]=1): s_t = sum_{i=1..L} C_i s_{t-i}.
inline int berlekamp_massey(const std::vector<uint8_t>& s, Poly& C) {
    const int n = (int)s.size();
    const int words = (n + 64) / 64 + 1;
    Poly Cc(words, 0), B(words, 0), T(words, 0), rev(words, 0);
    Cc[0] = 1; B[0] = 1;
    int L = 0, m = 1;
    for (int t = 0; t < n; ++t) {
        // rev bit j = s[t-j]
        for (int w = words - 1; w > 0; --w) rev[w] = (rev[w] << 1) | (rev[w - 1] >> 63);
        rev[0] = (rev[0] << 1) | (uint64_t)s[t];
        // discrepancy d = sum_{i=0..L} C_i s_{t-i}
        uint64_t acc = 0;
        const int lw = L / 64 + 1;
        for (int w = 0; w < lw; ++w) acc ^= Cc[w] & rev[w];
        const int d = __builtin_parityll(acc);
        if (d) {
            T = Cc;
            // C ^= B << m
            const int ws = m >> 6, bs = m & 63;
            for (int w = words - 1; w >= ws; --w) {
                uint64_t v = B[w - ws] << bs;
                if (bs && w - ws - 1 >= 0) v |= B[w - ws - 1] >> (64 - bs);
                Cc[w] ^= v;
            }
            if (2 * L <= t) { L = t + 1 - L; B = T; m = 1; }
            else ++m;
        } else ++m;
    }
    C = Cc;
    return L;
}

struct Field {
    Poly phi;                    // characteristic polynomial, bit DEG set
    std::vector<Poly> phi_sh;    // phi << r, r = 0..63 (PW+1 words)
    bool ok = false;

    void init() {
        // bit 0 of the raw stream of an arbitrary nonzero state
        const int n = 2 * DEG + 200;
        const std::vector<uint64_t> w = raw_stream(5489ULL, (size_t)n + 400);
        std::vector<uint8_t> s((size_t)n);
        for (int t = 0; t < n; ++t) s[t] = (uint8_t)(w[(size_t)t + 350] & 1ULL);
        Poly C;
        const int L = berlekamp_massey(s, C);
        ok = (L == DEG);
        phi.assign(PW + 1, 0);
        // phi(x) = x^L C(1/x): phi_k = C_{L-k}
        for (int k = 0; k <= L && k <= DEG; ++k) if (get_bit(C, L - k)) flip_bit(phi, k);
        phi_sh.assign(64, Poly(PW + 2, 0));
        for (int r = 0; r < 64; ++r)
            for (int wd = 0; wd <= PW; ++wd) {
                phi_sh[r][wd] ^= r ? (phi[wd] << r) : phi[wd];
                if (r) phi_sh[r][wd + 1] ^= phi[wd] >> (64 - r);
            }
    }
    // reduce a product (2*PW words) modulo phi, in place; result in the low PW words
    void reduce(Poly& p) const {
        for (int i = 2 * PW * 64 - 1; i >= DEG; --i) {
            if (!((p[(size_t)i >> 6] >> (i & 63)) & 1ULL)) continue;
            const int sh = i - DEG, ws = sh >> 6, r = sh & 63;
            const Poly& f = phi_sh[r];
            for (int wd = 0; wd < PW + 2 && ws + wd < 2 * PW; ++wd) p[(size_t)ws + wd] ^= f[wd];
        }
    }
    Poly mulmod(const Poly& a, const Poly& b) const {
        // 64 shifted copies of b
        std::vector<Poly> bs(64, Poly(PW + 1, 0));
        for (int r = 0; r < 64; ++r)
            for (int wd = 0; wd < PW; ++wd) {
                bs[r][wd] ^= r ? (b[wd] << r) : b[wd];
                if (r) bs[r][wd + 1] ^= b[wd] >> (64 - r);
            }
        Poly prod(2 * PW, 0);
        for (int i = 0; i < PW * 64; ++i) {
            if (!get_bit(a, i)) continue;
            const int ws = i >> 6, r = i & 63;
            const Poly& f = bs[r];
            for (int wd = 0; wd <= PW && ws + wd < 2 * PW; ++wd) prod[(size_t)ws + wd] ^= f[wd];
        }
        reduce(prod);
        prod.resize(PW);
        return prod;
    }
    Poly x_pow_2k(int k) const {  // x^(2^k) mod phi
        Poly p(PW, 0);
        flip_bit(p, 1);
        for (int i = 0; i < k; ++i) p = mulmod(p, p);
        return p;
    }
};

// W[t+d .. t+d+312) from W[t .. t+312+DEG) and g = x^d mod phi
inline void jump_window(const uint64_t* w_at_t, const Poly& g, uint64_t* out312) {
    for (int u = 0; u < 312; ++u) out312[u] = 0;
    for (int i = 0; i < DEG; ++i) {
        if (!get_bit(g, i)) continue;
        const uint64_t* src = w_at_t + i;
        for (int u = 0; u < 312; ++u) out312[u] ^= src[u];
    }
}

struct Table {
    bool ok = false;
    std::vector<uint64_t> polys;  // NSEG * PW words; entry c = x^(c*SEG) mod phi (entry 0 unused)
};

// builds the table (one to two seconds of host time, once per process) and checks one jump against
// the sequential recurrence
inline const Table& table() {
    static Table T;
    static bool built = false;
    if (built) return T;
    built = true;
    Field F;
    F.init();
    if (!F.ok) return T;
    const Poly gS = F.x_pow_2k(SEG_LOG2);
    T.polys.assign((size_t)NSEG * PW, 0);
    // g_c = g_S^c: the first NT entries one after another, then NT host threads step through c, c+NT, c+2NT, ...
    // with the common factor g_{NT*S}
    int NT = (int)std::thread::hardware_concurrency();
    NT = NT < 1 ? 1 : (NT > 16 ? 16 : NT);
    std::vector<Poly> first((size_t)NT + 1);
    first[1] = gS;
    for (int c = 2; c <= NT; ++c) first[c] = F.mulmod(first[c - 1], gS);
    const Poly step = first[NT];
    std::vector<std::thread> pool;
    for (int th = 1; th <= NT; ++th)
        pool.emplace_back([&, th]() {
            Poly cur = first[th];
            for (int c = th; c < NSEG; c += NT) {
                if (c > th) cur = F.mulmod(cur, step);
                std::memcpy(&T.polys[(size_t)c * PW], cur.data(), sizeof(uint64_t) * PW);
            }
        });
    for (auto& t : pool) t.join();
    // self test: jump by SEG and by 3*SEG from t = 1000 of seed 1
    const size_t need = 1000 + 3 * (size_t)SEG + 400 + LEAD;
    const std::vector<uint64_t> w = raw_stream(1ULL, need);
    bool good = true;
    for (int c : {1, 3}) {
        Poly g(PW);
        std::memcpy(g.data(), &T.polys[(size_t)c * PW], sizeof(uint64_t) * PW);
        uint64_t out[312];
        jump_window(&w[1000], g, out);
        for (int u = 0; u < 312; ++u) if (out[u] != w[1000 + (size_t)c * SEG + u]) good = false;
    }
    T.ok = good;
    return T;
}

}  // namespace mtjump
}  // namespace cbsg
