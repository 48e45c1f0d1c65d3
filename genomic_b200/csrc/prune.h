// prune.h -- undo.splits = "prune" (DNAcopy's changepoints.prune as ported in lib/cbs/CBS.cpp:229-320), host side.
// Off by default (undo_prune=false).  It is a search over subsets of the change points of ONE unit on per-segment sums --
// a few dozen numbers -- so it stays on the host, as in the reference.
//
// What the reference computes: with k change points, for r = k-1, k-2, .. 1 the r-subset with the smallest within-segment
// sum of squares W_r; as soon as W_r / W_k exceeds 1 + cutoff, the best (r+1)-subset is the answer (all change points if that
// happens at r = k-1, one segment if it never happens).  Parity needs its floating point and its enumeration quirks:
//   * a merged group's term is (s_a + s_{a+1} + .. + s_b, added left to right)^2 / count, terms are added left to right;
//   * subsets are visited in lexicographic order and `<=` lets a later subset replace an equal earlier one;
//   * the reference's stepping stops when the first kept change point reaches k - r, so that last subset is never looked at.
// Design here: the group terms are tabulated once (term[a][b] for all a <= b), and the subsets are walked depth first with the
// partial sum of the finished groups carried along -- the additions happen in the reference's order, so every W is the same
// double, but a subset costs O(1) per level instead of a pass over all segments.
#pragma once
#include <limits>
#include <vector>

namespace cbsg {

class PruneSearch {
public:
    PruneSearch(const double* x, int n, const std::vector<int>& lseg) : nseg_((int)lseg.size()), n_(n), lseg_(lseg) {
        total_sq_ = 0.0;
        for (int i = 0; i < n; ++i) total_sq_ += x[i] * x[i];
        std::vector<double> segsum((size_t)nseg_, 0.0);
        for (int s = 0, pos = 0; s < nseg_; ++s)
            for (int j = 0; j < lseg[s]; ++j) segsum[(size_t)s] += x[pos++];
        term_.assign((size_t)nseg_ * nseg_, 0.0);
        for (int a = 0; a < nseg_; ++a) {
            double run = 0.0;
            int cnt = 0;
            for (int b = a; b < nseg_; ++b) {
                run += segsum[(size_t)b];
                cnt += lseg[(size_t)b];
                term_[(size_t)a * nseg_ + b] = run * run / (double)cnt;
            }
        }
    }

    // smallest W over the r-subsets the reference visits; `cuts` receives the minimiser (cut c = boundary after segment c)
    double best_subset(int r, std::vector<int>& cuts) {
        r_ = r;
        last_first_ = (nseg_ - 1) - r;  // subsets whose first cut reaches this index are never visited
        best_ = std::numeric_limits<double>::infinity();
        cur_.assign((size_t)r, 0);
        walk(0, -1, 0.0);
        cuts = best_cuts_;
        return best_;
    }

    // W of the full model (every change point kept)
    double full_model() const {
        double acc = 0.0;
        for (int s = 0; s < nseg_; ++s) acc += term(s, s);
        return total_sq_ - acc;
    }

    std::vector<int> lengths_of(const std::vector<int>& cuts) const {
        std::vector<int> out;
        int seg = 0, taken = 0;
        for (int c : cuts) {
            int len = 0;
            for (; seg <= c; ++seg) len += lseg_[(size_t)seg];
            out.push_back(len);
            taken += len;
        }
        out.push_back(n_ - taken);
        return out;
    }

private:
    double term(int a, int b) const { return term_[(size_t)a * nseg_ + b]; }

    // depth = cuts already placed, prev = the last of them (-1: none), acc = sum of the terms of the groups they close
    void walk(int depth, int prev, double acc) {
        if (depth == r_) {
            const double w = total_sq_ - (acc + term(prev + 1, nseg_ - 1));
            if (w <= best_) { best_ = w; best_cuts_ = cur_; }
            return;
        }
        // cut `depth` may sit anywhere that leaves room for the r - depth - 1 cuts after it
        const int hi = (depth == 0) ? last_first_ - 1 : last_first_ + depth;
        for (int c = prev + 1; c <= hi; ++c) {
            cur_[(size_t)depth] = c;
            walk(depth + 1, c, acc + term(prev + 1, c));
        }
    }

    int nseg_, n_, r_ = 0, last_first_ = 0;
    const std::vector<int>& lseg_;
    double total_sq_ = 0.0, best_ = 0.0;
    std::vector<double> term_;
    std::vector<int> cur_, best_cuts_;
};

// x: the unit's values (as segmented), lseg: segment lengths; returns the pruned lengths
inline std::vector<int> prune_lengths(const double* x, int n, const std::vector<int>& lseg, double pcut) {
    const int k = (int)lseg.size() - 1;  // change points
    if (k <= 0) return lseg;
    PruneSearch search(x, n, lseg);
    const double w_full = search.full_model();
    std::vector<int> finer((size_t)k), cuts;
    for (int c = 0; c < k; ++c) finer[(size_t)c] = c;
    for (int r = k - 1; r >= 1; --r) {
        const double w = search.best_subset(r, cuts);
        if (w / w_full > 1.0 + pcut) return search.lengths_of(finer);  // dropping down to r change points costs too much
        finer = cuts;
    }
    return std::vector<int>{n};
}

}  // namespace cbsg
