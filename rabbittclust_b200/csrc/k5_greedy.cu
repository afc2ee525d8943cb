// k5_greedy.cu -- K5: greedy representative scan (clust-greedy), host driver + representative table.
//
// Replaces MinHashGreedyClusterWithInvertedIndex (reference src/greedy.cpp:986-1399).  The reference
// walks the genomes serially; genome j probes a dynamic hash -> representative-ids index, counts the
// shared hashes per touched representative (:1148-1164), filters by common_min (:1107-1113,
// :1211-1219), picks the best representative (:1198-1309) or becomes a new one (:1321-1336).
//
// B200 design (DESIGN.md "K5"): genomes are processed in batches.  For a batch the K3 tile kernel
// (candidate mode, first-hit tracking) computes  batch x current-representatives  and  batch x batch
// (earlier member only) and emits the sparse list of candidates (query, rep, common, first_hit) that
// can pass the common_min filter.  The host then replays the reference's serial decision per query in
// genome order -- candidates sorted in the reference's `touched` order: (index of the first shared
// hash in the query sketch, representative creation order) -- so the result is exactly the reference's
// with one thread.  Genomes that become representatives are appended to a device
// table kept in the same 32-interleaved layout as the main sketch set.
#include "tc_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>

namespace {

// gt[goff[g] + row*32 + lane] = rank `row` of sketch ids[g*32+lane] (TC_SENT_A padding)
__global__ void interleave_gather_kernel(const uint32_t *ranks, const unsigned long long *off, const uint32_t *ids,
                                         uint32_t nmembers, const unsigned long long *goff, uint32_t g0, uint32_t *gt)
{
    const uint32_t g = g0 + blockIdx.y;
    const unsigned long long rows = (goff[g + 1] - goff[g]) / TC_GROUP;
    const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * TC_GROUP) return;
    const uint32_t lane = (uint32_t)(e & 31), row = (uint32_t)(e >> 5);
    const uint32_t mem = g * TC_GROUP + lane;
    uint32_t v = TC_SENT_A;
    if (mem < nmembers) {
        const uint32_t id = ids[mem];
        const unsigned long long o = off[id], len = off[id + 1] - o;
        if (row < len) v = ranks[o + row];
    }
    gt[goff[g] + e] = v;
}

template <typename T>
int grow(T **p, uint64_t *cap, uint64_t need, uint64_t keep, cudaStream_t st)
{
    if (need <= *cap) return 0;
    uint64_t ncap = std::max<uint64_t>(need, *cap * 2);
    ncap = std::max<uint64_t>(ncap, 1024);
    T *q = nullptr;
    TC_CUDA(cudaMalloc(&q, ncap * sizeof(T)));
    if (*p && keep) TC_CUDA(cudaMemcpyAsync(q, *p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
    TC_CUDA(cudaStreamSynchronize(st));
    cudaFree(*p);
    *p = q; *cap = ncap;
    return 0;
}

}  // namespace

void tc_rep_table_free(tc_ctx *ctx)
{
    cudaFree(ctx->d_rep_gt); cudaFree(ctx->d_rep_goff); cudaFree(ctx->d_rep_len); cudaFree(ctx->d_rep_ids);
    cudaFree(ctx->d_cands);
    ctx->d_rep_gt = nullptr; ctx->d_rep_goff = nullptr; ctx->d_rep_len = nullptr; ctx->d_rep_ids = nullptr;
    ctx->d_cands = nullptr; ctx->cand_cap = 0;
    ctx->rep_gt_cap = ctx->rep_goff_cap = ctx->rep_len_cap = ctx->rep_ids_cap = 0;
    ctx->nrep = 0; ctx->h_rep_ids.clear(); ctx->h_rep_goff.assign(1, 0);
}

void tc_rep_table_reset(tc_ctx *ctx)
{
    ctx->nrep = 0; ctx->h_rep_ids.clear(); ctx->h_rep_goff.assign(1, 0);
}

// append representatives (global sketch ids, in creation order) to the device table
int tc_rep_table_append(tc_ctx *ctx, const uint32_t *ids, uint32_t count)
{
    if (!count) return 0;
    cudaStream_t st = ctx->stream;
    const uint32_t old_n = ctx->nrep, new_n = old_n + count;
    const uint32_t g_first = old_n / TC_GROUP;                       // first group whose content changes
    const uint32_t ng_new = (new_n + TC_GROUP - 1) / TC_GROUP;
    for (uint32_t i = 0; i < count; i++) {
        TC_CHECK(ids[i] < ctx->n, "rep table: sketch id %u out of range", ids[i]);
        ctx->h_rep_ids.push_back(ids[i]);
    }
    ctx->h_rep_goff.resize(ng_new + 1);
    uint64_t max_rows = 0;
    for (uint32_t g = g_first; g < ng_new; g++) {
        uint32_t rows = 0;
        for (uint32_t l = 0; l < TC_GROUP && g * TC_GROUP + l < new_n; l++)
            rows = std::max(rows, ctx->h_size[ctx->h_rep_ids[g * TC_GROUP + l]]);
        ctx->h_rep_goff[g + 1] = ctx->h_rep_goff[g] + (uint64_t)rows * TC_GROUP;
        max_rows = std::max<uint64_t>(max_rows, rows);
    }
    if (grow(&ctx->d_rep_gt, &ctx->rep_gt_cap, ctx->h_rep_goff[ng_new] + 1, ctx->h_rep_goff[g_first], st)) return 1;
    if (grow(&ctx->d_rep_goff, &ctx->rep_goff_cap, (uint64_t)ng_new + 1, (uint64_t)g_first + 1, st)) return 1;
    if (grow(&ctx->d_rep_len, &ctx->rep_len_cap, (uint64_t)ng_new * TC_GROUP, old_n, st)) return 1;
    if (grow(&ctx->d_rep_ids, &ctx->rep_ids_cap, (uint64_t)ng_new * TC_GROUP, old_n, st)) return 1;
    std::vector<uint32_t> len(count);
    for (uint32_t i = 0; i < count; i++) len[i] = ctx->h_size[ids[i]];
    TC_CUDA(cudaMemcpyAsync(ctx->d_rep_goff + g_first, ctx->h_rep_goff.data() + g_first,
                            (size_t)(ng_new + 1 - g_first) * 8, cudaMemcpyHostToDevice, st));
    TC_CUDA(cudaMemcpyAsync(ctx->d_rep_ids + old_n, ids, (size_t)count * 4, cudaMemcpyHostToDevice, st));
    TC_CUDA(cudaMemcpyAsync(ctx->d_rep_len + old_n, len.data(), (size_t)count * 4, cudaMemcpyHostToDevice, st));
    ctx->nrep = new_n;
    if (max_rows > 0) {
        dim3 grid((unsigned)((max_rows * TC_GROUP + 255) / 256), ng_new - g_first);
        interleave_gather_kernel<<<grid, 256, 0, st>>>(ctx->d_ranks, reinterpret_cast<unsigned long long *>(ctx->d_off),
                                                       ctx->d_rep_ids, new_n,
                                                       reinterpret_cast<unsigned long long *>(ctx->d_rep_goff), g_first,
                                                       ctx->d_rep_gt);
        ctx->launches++;
        TC_CUDA(cudaGetLastError());
    }
    // the host staging vectors (len) must outlive the async copies
    TC_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------ host formulas (exact)
extern "C" int tc_greedy_common_min(double threshold, int kmer_size, int size_ref, int size_qry, int containment)
{
    // greedy.cpp:1107-1113 (fixed size: size_ref == size_qry) and :1211-1219
    const double x = std::exp(-threshold * kmer_size);
    const double jaccard_min = x / (2.0 - x);
    if (containment) return (int)std::ceil(jaccard_min * std::min(size_ref, size_qry));
    return (int)std::ceil(jaccard_min * (size_ref + size_qry) / (1.0 + jaccard_min));
}

extern "C" double tc_greedy_dist(int common, int size_ref, int size_qry, int containment, int kmer_size)
{
    // greedy.cpp:1245-1274
    double jaccard;
    if (containment) {
        const int min_size = std::min(size_ref, size_qry);
        if (min_size == 0) return 1.0;
        jaccard = (double)common / min_size;
    } else {
        const int denom = size_ref + size_qry - common;
        if (denom == 0) return 0.0;
        jaccard = (double)common / denom;
    }
    if (jaccard >= 1.0) return 0.0;
    if (jaccard <= 0.0) return 1.0;
    double dist = -std::log(2.0 * jaccard / (1.0 + jaccard)) / kmer_size;
    if (dist > 1.0) dist = 1.0;
    return dist;
}

// ------------------------------------------------------------------------- validation entry
extern "C" int tc_greedy_scan(tc_ctx *ctx, uint32_t q0, uint32_t q1, const uint32_t *rep_ids, uint32_t nrep,
                              uint32_t *common, uint32_t *first_hit)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->n > 0, "tc_greedy_scan: no sketches uploaded");
    TC_CHECK(q0 <= q1 && q1 <= ctx->n, "tc_greedy_scan: query range out of bounds");
    const size_t cells = (size_t)(q1 - q0) * nrep;
    if (!cells) return 0;
    tc_rep_table_reset(ctx);
    if (tc_rep_table_append(ctx, rep_ids, nrep)) return 1;
    uint32_t *d_c = nullptr, *d_f = nullptr;
    TC_CUDA(cudaMalloc(&d_c, cells * 4));
    int rc = 1;
    do {
        if (cudaMalloc(&d_f, cells * 4) != cudaSuccess) { tc_set_error("tc_greedy_scan: out of device memory"); break; }
        if (tc_launch_tiles_rep_matrix(ctx, q0, q1, d_c, d_f)) break;
        if (cudaMemcpyAsync(common, d_c, cells * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(first_hit, d_f, cells * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            tc_set_error("tc_greedy_scan: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rc = 0;
    } while (0);
    cudaFree(d_c); cudaFree(d_f);
    return rc;
}

// Iteration order of the reference's candidate table (src/phmap.h, vendored parallel-hashmap in its deterministic build):
// `phmap::flat_hash_map<int,int> m; m.reserve(1000); m[key]++` for distinct keys in insertion order.  hash = high + low
// 64 bits of key * 0xde5fb9d2630458e9 (phmap_utils.h:67-77); slot search from (hash >> 7) & capacity over 16-wide groups
// with triangular steps, first empty position of the first group that has one (phmap.h:145-167, 2196-2207; the control
// array ends with a sentinel and 16 cloned bytes); capacity 2047 after reserve(1000) (:573-609, 1698-1714), growth to
// 2*capacity+1 once capacity - capacity/8 entries are in, re-inserting in slot order (:2054-2079, 2144-2154, 2252-2269).
// Checked against the real container in tests (ref_harness phorder -> oracle -> this).
struct PhmapOrder {
    std::vector<uint8_t> full; std::vector<int> key;
    uint64_t cap = 0; size_t size = 0, growth = 0;
    static uint64_t mix(uint64_t a) { const unsigned __int128 p = (unsigned __int128)a * 0xde5fb9d2630458e9ULL; return (uint64_t)(p >> 64) + (uint64_t)p; }
    size_t first_non_full(uint64_t h) const
    {
        uint64_t off = (h >> 7) & cap, idx = 0;
        for (;;) {
            for (uint64_t i = 0; i < 16; i++) {
                const uint64_t p = off + i;
                if (p == cap) continue;                                  // the sentinel
                if (!full[p < cap ? p : p - cap - 1]) return (size_t)((off + i) & cap);
            }
            idx += 16; off = (off + idx) & cap;
        }
    }
    void resize(uint64_t ncap)
    {
        std::vector<uint8_t> of; std::vector<int> ok;
        of.swap(full); ok.swap(key);
        const uint64_t oc = cap;
        cap = ncap; full.assign(ncap, 0); key.assign(ncap, 0);
        growth = (size_t)(ncap - ncap / 8) - size;
        for (uint64_t i = 0; i < oc; i++)
            if (of[i]) { const size_t t = first_non_full(mix((uint64_t)ok[i])); full[t] = 1; key[t] = ok[i]; }
    }
    // keys: distinct, in insertion order; returns them in iteration order
    std::vector<int> order(const std::vector<int> &keys)
    {
        cap = 0; size = 0; full.clear(); key.clear();
        resize(2047);
        for (int k : keys) {
            size_t t = first_non_full(mix((uint64_t)k));
            if (growth == 0) { resize(cap * 2 + 1); t = first_non_full(mix((uint64_t)k)); }
            size++; growth--;
            full[t] = 1; key[t] = k;
        }
        std::vector<int> out;
        out.reserve(keys.size());
        for (uint64_t i = 0; i < cap; i++) if (full[i]) out.push_back(key[i]);
        return out;
    }
};

// --------------------------------------------------------------------- full greedy clustering
// brute = false: MinHashGreedyClusterWithInvertedIndex (greedy.cpp:986-1399), semantic A, `touched` tie order.
// brute = true : greedyCluster (greedy.cpp:285-351): MinHash::distance (semantic B with sketchSize semb_size) or
//                containDistance of the query against every representative; nearest wins, ties -> earliest-created
//                representative (std::map<double,int>::insert keeps the first key, one thread).
// kssd = true  : KssdGreedyClusterWithInvertedIndex (greedy.cpp:566-899) with one thread: per-pair common_min,
//                best = largest Jaccard (strict >) in `touched` order, no distance test beyond common_min.
//
// Several GPUs ("sequential over genomes, parallel over representatives"): cs[0..G) are contexts that all
// hold the same sketch set; the representative table is dealt round-robin over them (representative number c lives on
// context c mod G), every context scans the batch against its share on its own host thread, the candidate lists are
// concatenated and the serial replay below runs unchanged -- it depends only on the union of the candidates, so the
// result is the single-GPU (= one-thread reference) result whatever G.
//
// incr = true  : KssdIncrementalCluster (greedy.cpp:1736-1880): the sketches [0, incr_start) are already clustered, their
//                representatives are seed[0..n_seed) in creation order; each later sketch joins the nearest
//                representative (two-sided size-ratio filter, floor()ed common filter, Mash distance <= threshold,
//                strictly smaller distance wins) or founds a cluster.  The reference walks its candidates in the
//                iteration order of its `phmap::flat_hash_map<int,int> candidate_counts` (reserve(1000)), so among EQUAL
//                distances the candidate that comes first in that table wins; the order is a pure function of the
//                touched representatives in first-touch order, reproduced by PhmapOrder below.  For that the device
//                emits every touched representative (common >= 1), not only those that can pass the filter.
//                kssd = false with incr: MinHashIncrementalCluster (greedy.cpp:1963-2130), the
//                same walk without the size-ratio filter and with the containment forms of the filter and the distance.
static int greedy_impl(tc_ctx *const *cs, int G, int kmer_size, int containment, double threshold, bool brute, uint32_t semb_size,
                       int32_t *rep_of, uint32_t *n_clusters, bool kssd = false, bool incr = false, uint32_t incr_start = 0,
                       const uint32_t *seed = nullptr, uint32_t n_seed = 0)
{
    tc_ctx *ctx = cs[0];
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->n > 0, "tc_greedy_cluster: no sketches uploaded");
    const uint32_t n = ctx->n;
    for (int r = 1; r < G; r++)
        TC_CHECK(cs[r] && cs[r]->n == n && cs[r]->total == ctx->total && cs[r]->h_size == ctx->h_size,
                 "tc_greedy_cluster_multi: context %d does not hold the same sketch set as context 0", r);
    // fixed-size fast path detection over the first 100 sketches (greedy.cpp:1092-1103)
    const int fixed_size = (int)ctx->h_size[0];
    bool all_fixed = !containment && !brute && !kssd && !incr;
    if (brute && containment) semb_size = 0;   // containDistance counts the full intersection
    for (uint32_t i = 1; i < std::min<uint32_t>(100, n) && all_fixed; i++)
        if ((int)ctx->h_size[i] != fixed_size) all_fixed = false;
    const int cmin_fixed = all_fixed ? tc_greedy_common_min(threshold, kmer_size, fixed_size, fixed_size, 0) : -1;
    const double x = std::exp(-threshold * kmer_size);
    const double jmin = x / (2.0 - x);
    const double max_ratio = 2.0 * std::exp(threshold * kmer_size) - 1.0;        // calculateMaxSizeRatio, greedy.cpp:162-173

    uint64_t cap = 1ull << 24;
    if (const char *s = getenv("TCLUST_CAND_CAP")) cap = std::max<uint64_t>(strtoull(s, nullptr, 10), 1024);
    for (int r = 0; r < G; r++) {
        TC_CUDA(cudaSetDevice(cs[r]->device));
        if (cs[r]->cand_cap < cap) {
            cudaFree(cs[r]->d_cands); cs[r]->d_cands = nullptr; cs[r]->cand_cap = 0;
            TC_CUDA(cudaMalloc(&cs[r]->d_cands, cap * sizeof(tc_cand)));
            cs[r]->cand_cap = cap;
        }
        tc_rep_table_reset(cs[r]);
    }
    uint32_t batch = 4096;
    if (const char *s = getenv("TCLUST_GREEDY_BATCH")) batch = (uint32_t)std::max<long>(atol(s), 32);
    batch = (batch + TC_GROUP - 1) / TC_GROUP * TC_GROUP;

    const char *gimpl = getenv("TCLUST_GREEDY_IMPL");
    const bool use_join = semb_size == 0 && !(gimpl && !strcmp(gimpl, "merge"));
    std::vector<int32_t> order_of(n, -1);      // representative creation order, -1 = not a representative
    std::vector<uint32_t> rep_by_order;        // creation number -> sketch id
    PhmapOrder phmap_order;
    std::vector<tc_cand> cands;
    std::vector<std::vector<tc_cand>> part_cands(G);
    std::vector<std::string> part_err(G);
    std::vector<char> part_over(G, 0);
    std::vector<unsigned long long> part_cnt(G, 0);
    // context r's share of one batch: batch x its representatives (+ on context 0: batch x batch, earlier member only)
    auto scan_part = [&](int r, uint32_t b0, uint32_t b1) {
        tc_ctx *c = cs[r];
        cudaStream_t st = c->stream;
        auto fail = [&](const char *what) { part_err[r] = std::string(what) + ": " + tc_last_error(); };
        if (cudaSetDevice(c->device) != cudaSuccess || cudaMemsetAsync(c->d_edge_count, 0, 2 * sizeof(unsigned long long), st) != cudaSuccess) {
            part_err[r] = "cudaSetDevice / memset failed"; return;
        }
        // candidates: the hash-join kernel (first_hit = smallest shared rank, same `touched` order), except for the
        // semantic-B counts of the brute-force rule, which need the ordered merge (TCLUST_GREEDY_IMPL=merge forces it)
        if (use_join) {
            // incremental rules: every touched representative (common >= 1), the tie order depends on all of them
            const int cmin_dev = incr ? 1 : cmin_fixed;
            if (c->nrep > 0 && tc_launch_join_cands(c, b0, b1, 1, cmin_dev, containment, jmin)) { fail("representative scan"); return; }
            if (r == 0 && tc_launch_join_cands(c, b0, b1, 0, cmin_dev, containment, jmin)) { fail("batch scan"); return; }
        } else {
            if (c->nrep > 0 && tc_launch_tiles_cands(c, b0, b1, 1, cmin_fixed, containment, jmin, 0, 1, semb_size)) { fail("representative scan"); return; }
            if (r == 0 && tc_launch_tiles_cands(c, b0, b1, 0, cmin_fixed, containment, jmin, 0, 1, semb_size)) { fail("batch scan"); return; }
        }
        unsigned long long cnt[2];
        if (cudaMemcpyAsync(cnt, c->d_edge_count, sizeof cnt, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            part_err[r] = std::string("candidate count: ") + cudaGetErrorString(cudaGetLastError()); return;
        }
        part_cnt[r] = cnt[0];
        part_over[r] = cnt[1] != 0 || cnt[0] > c->cand_cap;
        if (part_over[r]) return;
        part_cands[r].resize(cnt[0]);
        if (cnt[0] && (cudaMemcpyAsync(part_cands[r].data(), c->d_cands, cnt[0] * sizeof(tc_cand), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                       cudaStreamSynchronize(st) != cudaSuccess))
            part_err[r] = std::string("candidate download: ") + cudaGetErrorString(cudaGetLastError());
    };
    std::vector<uint32_t> start;
    std::vector<tc_cand> sorted;
    std::vector<uint32_t> new_reps;
    uint32_t nrep_total = 0;
    // creation number c of a representative lives on context c mod G
    auto append_reps = [&](const std::vector<uint32_t> &reps) -> int {
        if (G == 1) return tc_rep_table_append(ctx, reps.data(), (uint32_t)reps.size());
        std::vector<std::vector<uint32_t>> share(G);
        for (uint32_t q : reps) share[(uint32_t)order_of[q] % (uint32_t)G].push_back(q);
        std::vector<std::thread> th;
        for (int r = 0; r < G; r++)
            th.emplace_back([&, r]() {
                if (share[r].empty()) return;
                if (cudaSetDevice(cs[r]->device) != cudaSuccess || tc_rep_table_append(cs[r], share[r].data(), (uint32_t)share[r].size()))
                    part_err[r] = std::string("representative table: ") + tc_last_error();
            });
        for (std::thread &t : th) t.join();
        for (int r = 0; r < G; r++) TC_CHECK(part_err[r].empty(), "tc_greedy_cluster: GPU %d: %s", cs[r]->device, part_err[r].c_str());
        return 0;
    };
    if (incr) {
        TC_CHECK(incr_start <= n && n_seed <= incr_start && (seed || !n_seed), "tc_greedy_incremental_kssd: bad representative list");
        for (uint32_t c = 0; c < n_seed; c++) {
            TC_CHECK(seed[c] < incr_start && order_of[seed[c]] < 0, "tc_greedy_incremental_kssd: representative %u (sketch %u) is out of "
                     "range or listed twice", c, seed[c]);
            order_of[seed[c]] = (int32_t)c;
            rep_by_order.push_back(seed[c]);
        }
        nrep_total = n_seed;
        if (append_reps(std::vector<uint32_t>(seed, seed + n_seed))) return 1;
    }

    for (uint32_t b0 = incr ? incr_start : 0; b0 < n;) {
        const uint32_t b1 = std::min<uint64_t>(n, (uint64_t)b0 + batch);
        if (G == 1) scan_part(0, b0, b1);
        else {
            std::vector<std::thread> th;
            for (int r = 0; r < G; r++) th.emplace_back(scan_part, r, b0, b1);
            for (std::thread &t : th) t.join();
        }
        bool over = false;
        unsigned long long worst = 0;
        for (int r = 0; r < G; r++) {
            TC_CHECK(part_err[r].empty(), "tc_greedy_cluster: GPU %d: %s", cs[r]->device, part_err[r].c_str());
            over |= part_over[r] != 0;
            worst = std::max(worst, part_cnt[r]);
        }
        if (over) {
            TC_CHECK(batch > TC_GROUP, "tc_greedy_cluster: %llu candidates of one 32-genome batch exceed the buffer "
                     "(%llu); raise TCLUST_CAND_CAP", worst, (unsigned long long)ctx->cand_cap);
            batch = std::max<uint32_t>(TC_GROUP, batch / 2 / TC_GROUP * TC_GROUP);
            continue;                                           // redo this batch with a smaller one
        }
        if (G == 1) cands.swap(part_cands[0]);
        else {
            cands.clear();
            for (int r = 0; r < G; r++) cands.insert(cands.end(), part_cands[r].begin(), part_cands[r].end());
        }
        // bucket by query (counting sort)
        const uint32_t nb = b1 - b0;
        start.assign(nb + 1, 0);
        for (const tc_cand &c : cands) start[c.query - b0 + 1]++;
        for (uint32_t i = 0; i < nb; i++) start[i + 1] += start[i];
        sorted.resize(cands.size());
        {
            std::vector<uint32_t> fill(start.begin(), start.end() - 1);
            for (const tc_cand &c : cands) sorted[fill[c.query - b0]++] = c;
        }
        new_reps.clear();
        for (uint32_t q = b0; q < b1; q++) {
            tc_cand *lo = sorted.data() + start[q - b0], *hi = sorted.data() + start[q - b0 + 1];
            // keep candidates that are representatives by now; `touched` order = (first shared query hash,
            // representative creation order)  -- greedy.cpp:1148-1164
            tc_cand *end = std::remove_if(lo, hi, [&](const tc_cand &c) { return order_of[c.rep] < 0; });
            std::sort(lo, end, [&](const tc_cand &a, const tc_cand &b) {
                if (!brute && a.first_hit != b.first_hit) return a.first_hit < b.first_hit;
                return order_of[a.rep] < order_of[b.rep];
            });
            const int size_ref = (int)ctx->h_size[q];
            int best_common = -1, best_rep = -1, n_tied = 0;
            double best_dist = std::numeric_limits<double>::max(), best_jaccard = -1.0;
            for (tc_cand *c = lo; c < end; c++) {
                const int common = (int)c->common, size_qry = (int)ctx->h_size[c->rep];
                if (brute) {   // greedy.cpp:297-334: dist <= threshold, smallest distance, first inserted wins
                    const double dist = containment ? tc_contain_distance(common, size_qry, size_ref, kmer_size)
                                                    : tc_distance_b(common, size_qry, size_ref, (int)semb_size, kmer_size);
                    if (dist <= threshold && dist < best_dist) { best_dist = dist; best_rep = (int)c->rep; }
                    continue;
                }
                if (incr) {   // greedy.cpp:1812-1846 (KSSD), :2049-2094 (MinHash); sizeQry = the new sketch, sizeRef = the representative
                    const int sizeQry = size_ref, sizeRef = size_qry;
                    if (sizeRef == 0) continue;
                    if (kssd) {
                        const double ratio = (double)sizeQry / sizeRef;
                        if (ratio > max_ratio || ratio < 1.0 / max_ratio) continue;
                    }
                    const int min_common_needed = containment ? (int)(jmin * std::min(sizeQry, sizeRef))
                                                              : (int)(jmin * (sizeQry + sizeRef) / (1.0 + jmin));
                    if (common < min_common_needed) continue;
                    // calculate_mash_distance (:103-160) / the inline form of :2069-2085: the same expression
                    const double dist = tc_greedy_dist(common, sizeRef, sizeQry, containment, kmer_size);
                    if (dist <= threshold) {
                        if (dist < best_dist) { best_dist = dist; best_rep = (int)c->rep; n_tied = 1; }
                        else if (dist == best_dist) n_tied++;
                    }
                    continue;
                }
                const int common_min = all_fixed ? cmin_fixed
                                                 : tc_greedy_common_min(threshold, kmer_size, size_ref, size_qry, containment);
                if (common < common_min) continue;
                if (kssd) {   // greedy.cpp:783-794: compare Jaccard directly, larger wins, first in touched order on ties
                    const int denom = size_ref + size_qry - common;
                    const double jaccard = denom == 0 ? 1.0 : (double)common / denom;
                    if (jaccard > best_jaccard) { best_jaccard = jaccard; best_rep = (int)c->rep; }
                    continue;
                }
                if (all_fixed) {
                    if (common > best_common) { best_common = common; best_rep = (int)c->rep; }
                } else {
                    const double dist = tc_greedy_dist(common, size_ref, size_qry, containment, kmer_size);
                    if (dist <= threshold && dist < best_dist) { best_dist = dist; best_common = common; best_rep = (int)c->rep; }
                }
            }
            if (incr && n_tied > 1) {
                // several representatives at the smallest distance: the reference takes the one that comes first in the
                // iteration order of its candidate table, filled in first-touch order = the sorted order of lo..end
                std::vector<int> keys;
                keys.reserve(end - lo);
                for (tc_cand *c = lo; c < end; c++) keys.push_back(order_of[c->rep]);
                const std::vector<int> it = phmap_order.order(keys);
                std::vector<int> tied;                     // creation numbers of the tied candidates (a handful)
                for (tc_cand *c = lo; c < end; c++) {
                    const int common = (int)c->common, sizeQry = size_ref, sizeRef = (int)ctx->h_size[c->rep];
                    if (sizeRef == 0) continue;
                    if (kssd) { const double ratio = (double)sizeQry / sizeRef; if (ratio > max_ratio || ratio < 1.0 / max_ratio) continue; }
                    const int need = containment ? (int)(jmin * std::min(sizeQry, sizeRef)) : (int)(jmin * (sizeQry + sizeRef) / (1.0 + jmin));
                    if (common < need) continue;
                    if (tc_greedy_dist(common, sizeRef, sizeQry, containment, kmer_size) == best_dist) tied.push_back(order_of[c->rep]);
                }
                for (int idx : it)
                    if (std::find(tied.begin(), tied.end(), idx) != tied.end()) { best_rep = (int)rep_by_order[idx]; break; }
            }
            if (best_rep != -1) {
                rep_of[q] = best_rep;
            } else {
                rep_of[q] = (int32_t)q;
                order_of[q] = (int32_t)nrep_total++;
                rep_by_order.push_back(q);
                new_reps.push_back(q);
            }
        }
        // the new representatives join the tables
        if (append_reps(new_reps)) return 1;
        b0 = b1;
    }
    cudaSetDevice(ctx->device);
    if (n_clusters) *n_clusters = nrep_total;
    return 0;
}

extern "C" int tc_greedy_cluster(tc_ctx *ctx, int kmer_size, int containment, double threshold, int32_t *rep_of,
                                 uint32_t *n_clusters)
{
    return greedy_impl(&ctx, 1, kmer_size, containment, threshold, false, 0, rep_of, n_clusters);
}

extern "C" int tc_greedy_cluster_kssd(tc_ctx *ctx, int kmer_size, double threshold, int32_t *rep_of, uint32_t *n_clusters)
{
    return greedy_impl(&ctx, 1, kmer_size, 0, threshold, false, 0, rep_of, n_clusters, true);
}

extern "C" int tc_greedy_incremental_kssd(tc_ctx *ctx, int kmer_size, double threshold, uint32_t start_index, const uint32_t *rep_ids,
                                          uint32_t n_reps, int32_t *rep_of, uint32_t *n_clusters)
{
    TC_CHECK(rep_of, "tc_greedy_incremental_kssd: rep_of is null");
    return greedy_impl(&ctx, 1, kmer_size, 0, threshold, false, 0, rep_of, n_clusters, true, true, start_index, rep_ids, n_reps);
}

extern "C" int tc_greedy_incremental(tc_ctx *ctx, int kmer_size, int containment, double threshold, uint32_t start_index,
                                     const uint32_t *rep_ids, uint32_t n_reps, int32_t *rep_of, uint32_t *n_clusters)
{
    TC_CHECK(rep_of, "tc_greedy_incremental: rep_of is null");
    TC_CHECK(containment == 0 || containment == 1, "tc_greedy_incremental: containment must be 0 or 1");
    return greedy_impl(&ctx, 1, kmer_size, containment, threshold, false, 0, rep_of, n_clusters, false, true, start_index, rep_ids, n_reps);
}

extern "C" int tc_greedy_cluster_multi(tc_ctx *const *ctxs, int nctx, int kmer_size, int containment, double threshold, int kssd,
                                       int32_t *rep_of, uint32_t *n_clusters)
{
    TC_CHECK(ctxs && nctx >= 1 && nctx <= 64, "tc_greedy_cluster_multi: bad context list");
    for (int r = 0; r < nctx; r++) TC_CHECK(ctxs[r], "tc_greedy_cluster_multi: context %d is null", r);
    return greedy_impl(ctxs, nctx, kmer_size, kssd ? 0 : containment, threshold, false, 0, rep_of, n_clusters, kssd != 0);
}

extern "C" int tc_greedy_cluster_bruteforce(tc_ctx *ctx, int kmer_size, uint32_t sketch_size, int containment,
                                            double threshold, int32_t *rep_of, uint32_t *n_clusters)
{
    TC_CHECK(containment || sketch_size > 0, "tc_greedy_cluster_bruteforce: sketch_size == 0");
    TC_CHECK(threshold < 1.0, "tc_greedy_cluster_bruteforce: threshold must be < 1 (pairs without a shared hash "
             "have distance 1 and are never candidates)");
    return greedy_impl(&ctx, 1, kmer_size, containment, threshold, true, sketch_size, rep_of, n_clusters);
}
