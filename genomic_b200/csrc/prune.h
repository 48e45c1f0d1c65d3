// prune.h -- undo.splits = "prune" (DNAcopy), host side: restates prune_segments / errssq /
// next_combination of lib/cbs/CBS.cpp:229-320.  Off by default (undo_prune=false); it is an exhaustive
// search over subsets of change points on per-segment sums, negligible next to the permutation tests,
// and therefore stays on the host.
#pragma once
#include <vector>

namespace cbsg {

// sum over merged groups of (sum x)^2 / count, groups delimited by the kept change points `keep[0..k)`
inline double grouped_ssq(const std::vector<int>& lseg, const std::vector<double>& segsum, const std::vector<int>& keep, int k) {
    const int nseg = (int)lseg.size();
    double out = 0.0;
    int from = 0;
    for (int part = 0; part <= k; ++part) {
        const int to = (part < k) ? keep[part] : nseg - 1;
        double s = 0.0;
        int cnt = 0;
        for (int i = from; i <= to; ++i) { s += segsum[i]; cnt += lseg[i]; }
        out += s * s / (double)cnt;
        from = to + 1;
    }
    return out;
}

// advance `keep` to the next r-subset in lexicographic order; false when exhausted (CBS.cpp:257-264)
inline bool next_subset(std::vector<int>& keep, int r, int nmr) {
    int i = r - 1;
    while (i >= 0 && keep[i] == nmr + i) --i;
    if (i < 0) return false;
    ++keep[i];
    for (int j = i + 1; j < r; ++j) keep[j] = keep[j - 1] + 1;
    return keep[0] != nmr;
}

// x: the unit's values (as segmented), lseg: segment lengths; returns the pruned lengths
inline std::vector<int> prune_lengths(const double* x, int n, const std::vector<int>& lseg, double pcut) {
    const int nseg = (int)lseg.size();
    if (nseg <= 1) return lseg;
    double ssq = 0.0;
    for (int i = 0; i < n; ++i) ssq += x[i] * x[i];
    std::vector<double> segsum((size_t)nseg, 0.0);
    for (int i = 0, pos = 0; i < nseg; ++i)
        for (int j = 0; j < lseg[i]; ++j) segsum[i] += x[pos++];
    const int k = nseg - 1;
    std::vector<int> keep((size_t)k), best_prev((size_t)k), best_cur((size_t)k);
    for (int i = 0; i < k; ++i) { keep[i] = i; best_prev[i] = i; }
    const double wssqk = ssq - grouped_ssq(lseg, segsum, keep, k);
    for (int j = k - 1; j >= 1; --j) {
        const int kmj = k - j;
        for (int i = 0; i < j; ++i) { keep[i] = i; best_cur[i] = i; }
        double wssqj = ssq - grouped_ssq(lseg, segsum, keep, j);
        while (next_subset(keep, j, kmj)) {
            const double w = ssq - grouped_ssq(lseg, segsum, keep, j);
            if (w <= wssqj) { wssqj = w; for (int i = 0; i < j; ++i) best_cur[i] = keep[i]; }
        }
        if (wssqj / wssqk > 1.0 + pcut) {  // the finer level (j+1 change points) is kept
            std::vector<int> cums((size_t)nseg);
            for (int i = 0, s = 0; i < nseg; ++i) { s += lseg[i]; cums[i] = s; }
            std::vector<int> out;
            int prev = 0;
            for (int i = 0; i <= j; ++i) { out.push_back(cums[best_prev[i]] - prev); prev = cums[best_prev[i]]; }
            out.push_back(n - prev);
            return out;
        }
        for (int i = 0; i < j; ++i) best_prev[i] = best_cur[i];
    }
    return std::vector<int>{n};
}

}  // namespace cbsg
