/* TEST INFRASTRUCTURE -- CPU restatement of cngpld::summarize_cn for ONE sample and chromosome
 * (/root/reference lib/cngpld/summarize.cpp:24-36 default positions, :41-75 value at a position, :77-100 the loop).
 * Pinned by the reference's golden vectors tests/data/cngpld_case{1..4}_* (tests/test_oracle.py); the reference
 * function itself cannot be compiled here (SegmentedSampleSet.hpp needs Boost and a generated config.h).
 * Only tests/, smoke() and bench.py's cpu_baseline leg may use anything under oracle/. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int cmp_u64(const void* a, const void* b) {
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}

/* summarize.cpp:24-36: sorted unique segment starts and ends; out has room for 2*n; returns the count */
long long orc_cn_default_positions(const uint64_t* start, const uint64_t* end, long long n, uint64_t* out) {
    long long m = 0, k = 0;
    for (long long i = 0; i < n; ++i) { out[m++] = start[i]; out[m++] = end[i]; }
    qsort(out, (size_t)m, sizeof(uint64_t), cmp_u64);
    for (long long i = 0; i < m; ++i) if (i == 0 || out[i] != out[k - 1]) out[k++] = out[i];
    return k;
}

/* summarize.cpp:41-75; returns -1 (invalid_argument) for a bad direction or a segment with start > end */
int orc_cn_at_position(const uint64_t* start, const uint64_t* end, const float* value, long long n, uint64_t pos,
                       int direction, double cutoff, double* out) {
    if (direction != 1 && direction != -1) return -1;
    long long overlap = 0, altered = 0;
    double sum = 0.0;
    for (long long i = 0; i < n; ++i) {
        if (start[i] > end[i]) return -1;
        if (start[i] <= pos && pos <= end[i]) {
            ++overlap;
            const double adj = (double)direction * (double)value[i];
            if (adj > cutoff) { sum += exp(adj); ++altered; }
        }
    }
    *out = altered == 0 ? 0.0 : sum / (double)overlap;
    return 0;
}

/* summarize.cpp:77-100; positions == NULL -> default positions; out_pos/out_value have room for max(2*n, n_pos) */
long long orc_summarize_cn(const uint64_t* start, const uint64_t* end, const float* value, long long n, int direction,
                           double cutoff, const uint64_t* positions, long long n_pos, uint64_t* out_pos, double* out_value) {
    if (!positions) n_pos = orc_cn_default_positions(start, end, n, out_pos);
    else for (long long i = 0; i < n_pos; ++i) out_pos[i] = positions[i];
    for (long long i = 0; i < n_pos; ++i)
        if (orc_cn_at_position(start, end, value, n, out_pos[i], direction, cutoff, out_value + i)) return -1;
    return n_pos;
}
