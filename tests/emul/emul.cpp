// TEST INFRASTRUCTURE: host build of the product's scheduler (cbs_core.h) and
// thread-per-permutation code (cbs_threads.h).  The kernels that only exist as CUDA
// (prep, scan, generator, count) are replaced by sequential stand-ins; the scan
// stand-in calls the oracle.  This checks the worklist state machine, RNG cursor
// bookkeeping, batching/early-exit logic and the shuffle/edge threads on CPU.
// It is NOT a product path and is never loaded by genomic_b200.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../genomic_b200/csrc/cbs_threads.h"
#include "../../oracle/cbs_oracle.h"

using namespace cbsg;

namespace {

void prep_seq(Dev& D, Task& t) {
    const long long base = D.unit_off[t.unit] + t.lo;
    const double* x = D.x + base;
    double* cur = D.cur + base;
    const int n = t.n, nb = t.nb;
    int flat = 1;
    for (int i = 0; i < n; ++i) if (!(fabs(x[i] - x[0]) < 1e-12)) { flat = 0; break; }
    t.alleq = flat;
    if (flat) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i];
    const double avg = s / (double)n;
    double tss = 0.0;
    for (int i = 0; i < n; ++i) cur[i] = x[i] - avg;
    for (int i = 0; i < n; ++i) tss += cur[i] * cur[i];
    t.tss = tss;
    int* bb = D.bbtab + base;
    for (int b = 0; b <= nb; ++b) bb[b] = block_end(n, nb, b);
    const double rn = (double)n;
    for (int L = 1; L < n; ++L) {
        const double rr = (double)L;
        D.factab[base + L] = rn / (rr * (rn - rr));
        D.gtab[base + L] = sqrt((rr * (rn - rr)) / rn);
    }
    PlainGet g{cur};
    prefix_and_block_stats(g, n, nb, bb, D.arena + t.off_sx, BlockStats(D.arena + t.off_bs, nb));
}

}  // namespace

extern "C" int64_t emul_segment_units(const double* values, const int64_t* unit_off, const uint64_t* unit_ids,
                                      int n_units, double alpha, int nperm, int min_width, int rng_mode, int chain,
                                      uint64_t seed, int first_batch, int max_batch, long long arena_cap,
                                      long long draws_cap, int max_live, int64_t cap, int* seg_count, int* lengths,
                                      double* means, uint64_t* draws_out, int* n_rounds, int split_cap,
                                      SplitRec* splits_out, int* n_splits_out) {
    const long long N = unit_off[n_units];
    Dev D;
    memset(&D, 0, sizeof(D));
    std::vector<long long> uoff(unit_off, unit_off + n_units + 1);
    D.x = values; D.unit_off = uoff.data(); D.unit_ids = unit_ids; D.n_units = n_units;
    D.prm.alpha = alpha; D.prm.nperm = nperm; D.prm.min_width = min_width; D.prm.rng_mode = rng_mode;
    D.prm.chain = chain; D.prm.seed = seed; D.prm.first_batch = first_batch; D.prm.max_batch = max_batch;
    D.prm.record_splits = split_cap > 0;
    std::vector<double> cur(N + 1), gtab(N + 1), factab(N + 1);
    std::vector<int> bbtab(N + 1);
    D.cur = cur.data(); D.gtab = gtab.data(); D.factab = factab.data(); D.bbtab = bbtab.data();
    D.task_cap = 4096;
    std::vector<Task> tasks(D.task_cap);
    std::vector<int> ring(D.task_cap);
    for (int i = 0; i < D.task_cap; ++i) ring[i] = i;
    D.tasks = tasks.data(); D.free_ring = ring.data(); D.free_head = 0; D.free_tail = D.task_cap;
    D.list_cap = 4 * D.task_cap + n_units + 16;
    std::vector<int> l0(D.list_cap), l1(D.list_cap);
    D.active[0] = l0.data(); D.active[1] = l1.data();
    D.n_chains = (rng_mode == RNG_MT) ? (chain ? 1 : n_units) : 0;
    std::vector<Chain> chains(D.n_chains > 0 ? D.n_chains : 1);
    for (int c = 0; c < D.n_chains; ++c) {
        Chain& ch = chains[c];
        memset(&ch, 0, sizeof(ch));
        ch.top = -1; ch.unit_cur = -1;
        ch.unit_next = chain ? 0 : c; ch.unit_last = chain ? n_units : c + 1;
        ch.need_off = -1; ch.prev_off = -1;
        mt_seed_next312(seed, ch.hist);
    }
    D.chains = chains.data();
    D.max_live = max_live;
    D.seg_cap = (int)cap;
    std::vector<SegRec> segs(D.seg_cap > 0 ? D.seg_cap : 1);
    D.segs = segs.data();
    D.split_cap = split_cap;
    std::vector<SplitRec> splits(split_cap > 0 ? split_cap : 1);
    D.splits = splits.data();
    std::vector<uint64_t> udraws(n_units + 1);
    D.unit_draws = udraws.data();
    D.arena_cap = arena_cap;
    std::vector<double> arena(arena_cap);
    D.arena = arena.data();
    D.rej_cap = 1 << 20;
    std::vector<int> rej(D.rej_cap);
    D.rej = rej.data();
    D.draws_cap = (rng_mode == RNG_MT) ? draws_cap : 0;  // as the product does
    std::vector<uint64_t> d0(draws_cap > 0 ? draws_cap : 1), d1(draws_cap > 0 ? draws_cap : 1);
    D.draws[0] = d0.data(); D.draws[1] = d1.data();
    // shared stream exactly as the product configures it: MT replay with one engine per unit
    D.shared_stream = (rng_mode == RNG_MT && !chain) ? 1 : 0;
    // ring of 2^k words + mirror; EMUL_STREAM_RING (words) shrinks the ring to exercise the window logic of the scheduler
    long long nmax_unit = 1;
    for (int u = 0; u < n_units; ++u) if (unit_off[u + 1] - unit_off[u] > nmax_unit) nmax_unit = unit_off[u + 1] - unit_off[u];
    const long long stream_ring_words = getenv("EMUL_STREAM_RING") ? atoll(getenv("EMUL_STREAM_RING")) : (1LL << 24);
    const long long stream_mirror_words = nmax_unit + 1024;
    const long long sring = D.shared_stream ? stream_ring_words : 0, smirror = D.shared_stream ? stream_mirror_words : 0;
    std::vector<uint64_t> stream(D.shared_stream ? (size_t)(sring + smirror) : 1);
    D.stream = stream.data(); D.stream_cap = sring; D.stream_mask = sring - 1; D.stream_mirror = smirror; D.stream_lo = 0;
    if (D.shared_stream) {
        uint64_t next[312];
        mt_seed_next312(seed, next);
        for (int u = 0; u < 312; ++u) stream_put(D, u, next[u]);
        D.stream_len = 312; D.stream_target = 312;
    }
    D.span_max = 1LL << 40;
    std::vector<int> shuf_store(3 * SHUF_NCLS * (size_t)(4 * 4096 + n_units + 32));
    for (int k = 0; k < SHUF_NCLS; ++k) D.shuf_p0[k] = shuf_store.data() + (size_t)(2 * SHUF_NCLS + k) * (4 * 4096 + n_units + 32);
    for (int k = 0; k < SHUF_NCLS; ++k) {
        D.shuf_item[k] = shuf_store.data() + (size_t)(2 * k) * (4 * 4096 + n_units + 32);
        D.shuf_prefix[k] = shuf_store.data() + (size_t)(2 * k + 1) * (4 * 4096 + n_units + 32);
    }
    std::vector<int> item_uprefix(D.list_cap + 1);
    D.item_uprefix = item_uprefix.data();
    std::vector<int> prep_task(D.list_cap), edgeprep_task(D.list_cap), item_prefix(D.list_cap + 1),
        edge_prefix(D.list_cap + 1), gen_chain(D.n_chains + 1);
    std::vector<PermItem> items(D.list_cap);
    std::vector<EdgeItem> edges(D.list_cap);
    D.prep_task = prep_task.data(); D.edgeprep_task = edgeprep_task.data(); D.item_prefix = item_prefix.data();
    D.edge_prefix = edge_prefix.data(); D.gen_chain = gen_chain.data(); D.items = items.data(); D.edges = edges.data();

    std::vector<double> px;
    int rounds = 0;
    long long st_sparse = 0, st_general = 0, st_defer = 0, st_items = 0;
    while (!D.done) {
        // count phase on last round's items
        for (int k = 0; k < D.n_items; ++k) count_item_seq(D, D.items[k]);
        Sched S(D);
        S.run_round();
        ++rounds;
        if (D.done) break;
        if (rounds > 1000000) { D.error = ERR_INTERNAL; break; }
        const int par = D.round & 1;
        if (D.shared_stream) mt_extend_stream_seq(D);
        for (int g = 0; g < D.n_gen; ++g) {
            Chain& ch = D.chains[D.gen_chain[g]];
            mt_generate_seq(ch, D.draws[par ^ 1], D.draws[par]);
        }
        for (int k = 0; k < D.n_prep; ++k) prep_seq(D, D.tasks[D.prep_task[k]]);
        for (int k = 0; k < D.n_items; ++k) {
            const PermItem& it = D.items[k];
            Task& t = D.tasks[it.task];
            const double* curp = D.cur + D.unit_off[t.unit] + t.lo;
            if (it.obs) {
                if (t.alleq) continue;
                const orc_tmax r = orc_tmaxo(curp, t.n, t.tss, D.prm.min_width, 0);
                t.ostat = r.stat; t.tmaxi = r.start + 1; t.tmaxj = r.end + 1;
                continue;
            }
            px.resize(t.n);
            for (int p = 0; p < it.P; ++p) {
                // the emulation always uses the global-memory thread (class 4 path): give it a scratch
                std::vector<double> A((size_t)t.n * it.P);
                Task tt = t; tt.off_A = 0;
                Dev DD = D; DD.arena = nullptr;
                {
                    const long long base = D.unit_off[t.unit] + t.lo;
                    DrawSrc src;
                    if (D.prm.rng_mode == RNG_MT) src.init_mt(draw_window(D, t.off_draw + (long long)p * t.n));
                    else src.init_philox(t.key, 0u, (uint32_t)(t.perms_done + p));
                    fy_shuffle_column(A.data(), it.P, p, D.cur + base, t.n, src);
                    ColumnGet g{A.data(), it.P, p};
                    prefix_and_block_stats(g, t.n, t.nb, D.bbtab + base, D.arena + t.off_sx + (long long)p * Sched::sx_stride(t.n),
                                           BlockStats(D.arena + t.off_bs + (long long)p * Sched::bs_stride(t.nb), t.nb));
                }
                for (int i = 0; i < t.n; ++i) px[i] = A[(size_t)i * it.P + p];
                // check the prefix sums the thread produced
                const double* sx = D.arena + t.off_sx + (long long)p * Sched::sx_stride(t.n);
                double run = 0.0;
                for (int i = 0; i < t.n; ++i) { run += px[i]; if (sx[i + 1] != run) { D.error = ERR_INTERNAL; } }
                const double pstat = orc_tmaxp(px.data(), t.n, t.tss, D.prm.min_width, 0);
                D.rej[t.off_rej + p] = (t.ostat * 0.99999 <= pstat) ? 1 : 0;
            }
        }
        for (int k = 0; k < D.n_edgeprep; ++k) {
            Task& t = D.tasks[D.edgeprep_task[k]];
            edgeprep_seq(D, t, 0);
            edgeprep_seq(D, t, 1);
        }
        for (int k = 0; k < D.n_edge; ++k) {
            const EdgeItem& e = D.edges[k];
            Task& t = D.tasks[e.task];
            const int threads = D.edge_prefix[k + 1] - D.edge_prefix[k];
            (e.sparse ? st_sparse : st_general) += 1;
            for (int th = 0; th < threads; ++th)
                t.e_nrej[e.side] += e.sparse ? edge_sparse_thread(D, t, e, th) : edge_general_thread(D, t, e, th);
        }
    }
    if (getenv("EMUL_STATS")) fprintf(stderr, "emul: rounds=%d sparse_items=%lld general_items=%lld perms=%llu\n", rounds, st_sparse, st_general, D.stat_perms);
    if (n_rounds) *n_rounds = rounds;
    if (D.error) return -(int64_t)D.error;
    // assemble per unit, sorted by lo; means = sequential sum / len (CBS.cpp:1014-1022)
    std::vector<std::vector<SegRec>> per(n_units);
    for (int k = 0; k < D.n_segs; ++k) per[D.segs[k].unit].push_back(D.segs[k]);
    int64_t total = 0;
    for (int u = 0; u < n_units; ++u) {
        auto& v = per[u];
        for (size_t a = 1; a < v.size(); ++a) for (size_t b = a; b > 0 && v[b].lo < v[b - 1].lo; --b) std::swap(v[b], v[b - 1]);
        seg_count[u] = (int)v.size();
        for (auto& s : v) {
            if (total >= cap) return -1;
            double acc = 0.0;
            for (int i = s.lo; i < s.hi; ++i) acc += values[unit_off[u] + i];
            lengths[total] = s.hi - s.lo;
            means[total] = acc / (double)(s.hi - s.lo);
            ++total;
        }
        if (draws_out) draws_out[u] = D.unit_draws[u];
    }
    if (splits_out) { for (int k = 0; k < D.n_splits && k < split_cap; ++k) splits_out[k] = D.splits[k]; }
    if (n_splits_out) *n_splits_out = D.n_splits;
    return total;
}
