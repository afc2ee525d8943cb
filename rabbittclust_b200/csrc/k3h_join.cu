// k3h_join.cu -- K3H: all-pairs sketch intersection as a tile-level hash join (HP-2, second
// generation of the dense pass; same results as k3_tile.cu, several times fewer instructions per pair).
//
// Replaces the same reference code as K3 (candidate loop of compute_minhash_mst, src/MST.cpp:1409-1435,
// 1467-1533, and the pairwise merges of RabbitSketch/src/MinHash.cpp:1199-1643): for every pair (i, j)
// the exact |A_i n A_j| (semantic A).  The reference's own default path is an inverted index
// (hash -> genome ids, src/SketchInfo.h:95-161); this kernel is the dense, data-independent GPU form of
// that idea: an inverted index of only 32 sketches at a time, held in shared memory.
//
// Design (DESIGN.md "K3H"):
//  * a work unit is a row block of 32 sketches x a segment of up to 4096 column sketches.  One persistent
//    CTA (1024 threads) per SM takes units from an atomic counter; unit u belongs to rank u % nparts.
//  * per unit the CTA builds, in shared memory, an open-addressing table  rank -> 32-bit mask of the row
//    sketches that contain it  (16384 slots, at most 10688 DISTINCT ranks per build) plus a 64 KB blocked-Bloom
//    presence filter (2 bits per key in one word, stored inverted).  A row block with more distinct ranks is cut
//    into P value ranges ("partitions"); k3h_distinct_kernel sizes P and the pivots from the distinct count once
//    per upload (32 related sketches share most ranks: P = 1), k3h_pivot_kernel is the fall-back for row blocks
//    too large to sort in shared memory.
//  * the column sketches stream through the table: one warp per column, 128 elements per step (4 coalesced
//    128-byte loads from the row-major u32 ranks, next step prefetched); a lane whose element misses the
//    presence filter is done after one LDS; hits fetch the mask from the table.  Lane r accumulates
//    count(row r, column): masks are redistributed with shuffles (few hits) or 32 ballots (many hits).
//  * between partitions the per-column state (32 counts + stream position; candidate mode: + 32 first shared
//    ranks) lives in a per-CTA scratch in global memory (L2 resident); after the last partition the epilogue
//    emits the candidate edge / matrix cell / greedy candidate.
//  * candidate mode (greedy scan, K5): columns are streamed through an id indirection (the representative
//    table), rows are restricted to the query batch.
//  Units are ordered segment-major, so at any moment all CTAs stream the same 16 MB of column data: the
//  column stream is served by L2, not HBM.
#include "tc_common.cuh"
#include <algorithm>

namespace {

constexpr int KH_THREADS = 1024;
constexpr uint32_t KH_SLOTS = 16384;          // table slots (key, mask), 128 KB
constexpr uint32_t KH_KCAP = 10752;           // max elements of the row block per partition (32 x 1000 -> 3 partitions, load 0.65)
constexpr uint32_t KH_FBITS = 19;             // presence filter: 2^19 bits = 64 KB, 2 bits per key in one word
constexpr uint32_t KH_FWORDS = (1u << KH_FBITS) / 32;
constexpr uint32_t KH_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t KH_CSEG = 4096;            // columns per unit
constexpr uint32_t KH_STATE = 33;             // words of state per column: 32 counts + position
constexpr uint32_t KH_STATE_CAND = 65;        // candidate mode: + 32 smallest shared ranks
constexpr size_t KH_SMEM = (size_t)KH_SLOTS * 8 + (size_t)KH_FWORDS * 4;
constexpr uint32_t KH_SORT_CAP = 32768;       // row blocks up to this many elements get distinct-count partitions
constexpr uint32_t KH_DREC = 8;               // per row block: distinct count, pivot candidates for P = 2, 3, 4, shared elements

enum { KH_MODE_MATRIX = 0, KH_MODE_EDGES = 1, KH_MODE_CAND = 2 };

struct KHArgs {
    const uint32_t *ranks;
    const unsigned long long *off;
    const uint32_t *size;
    uint32_t n;
    const uint32_t *pivots;         // interior pivots of row block rb: pivots[pivoff[rb] .. pivoff[rb+1])
    const uint32_t *pivoff;
    const tc_kh_unit *units;
    uint32_t nunits, part, nparts;
    unsigned long long *counters;   // [0] edge count, [1] overflow, [2] unit counter
    uint32_t *state;                // gridDim.x * KH_CSEG * KH_STATE words
    int mode;
    uint32_t *out; uint32_t i0, i1, j0, j1;
    tc_pair *edges; unsigned long long edge_cap; int radio;
    uint32_t row_min;               // edge mode: only rows i >= row_min (the start_index of the append flows)
    // candidate mode (greedy representative scan, K5): columns are the sketches col_ids[c] (nullptr: c itself) for c in
    // the unit's range, rows the queries [q0, q1); emits (query, rep, common, smallest shared rank) when common >= 1 and
    // common + 1 >= common_min(pair)  -- the rule of k3_tile.cu's candidate mode
    const uint32_t *col_ids; uint32_t q0, q1;
    tc_cand *cands; int lower_only, cmin_fixed, containment; double jmin, one_plus_jmin;
    // block join (K3J): union lists of the 32-sketch column blocks
    const uint32_t *bu_keys, *bu_mask, *bu_count;
    const uint32_t *multi;          // bit r = rank r occurs in at least two sketches (only those can be shared)
    // tile candidacy (k3t_tileflags.cu): bit cb of row rb set = blocks rb and cb share a rank; nullptr = visit every tile
    const uint32_t *tile_flags; uint32_t tile_words;
    // threshold-only clustering (tc_cluster_threshold): instead of appending a candidate edge, union (i, j) in a
    // device-resident union-find when J = common/denom >= uf_jmin (<=> reference distance <= threshold);
    // uf_perm: ids of the reordered copy -> input ids (nullptr: identity); distance mode in `containment`
    uint32_t *uf; const uint32_t *uf_perm; double uf_jmin;
};

// lock-free union-find (hook the larger root under the smaller: the root of a component is its smallest vertex)
__device__ __forceinline__ uint32_t kh_uf_find(uint32_t *p, uint32_t x)
{
    for (;;) {
        const uint32_t q = *(volatile uint32_t *)&p[x];
        if (q == x) return x;
        x = q;
    }
}
__device__ __forceinline__ void kh_uf_union(uint32_t *p, uint32_t a, uint32_t b)
{
    for (;;) {
        a = kh_uf_find(p, a); b = kh_uf_find(p, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        if (atomicCAS(&p[a], a, b) == a) return;
    }
}
// the pair (i, j) with `common` shared hashes joins one cluster iff its distance is within the threshold
__device__ __forceinline__ void kh_uf_edge(const KHArgs &a, uint32_t i, uint32_t j, uint32_t common, uint32_t s0, uint32_t s1)
{
    const uint32_t denom = a.containment ? min(s0, s1) : s0 + s1 - common;
    const double J = denom == 0 ? 0.0 : __ddiv_rn((double)common, (double)denom);
    if (!(J >= a.uf_jmin)) return;
    if (a.uf_perm) { i = __ldg(a.uf_perm + i); j = __ldg(a.uf_perm + j); }
    kh_uf_union(a.uf, i, j);
}

// presence filter: word = top (KH_FBITS-5) bits of h = x * golden, the two bit positions = low 10 bits of h
__device__ __forceinline__ uint32_t kh_fword(uint32_t h) { return h >> (32 - (KH_FBITS - 5)); }
// two bit positions from the low 10 bits of h; the funnel shift takes its amount mod 32 (one SHF.L.W each)
__device__ __forceinline__ uint32_t kh_fmask(uint32_t h) { return __funnelshift_l(0u, 1u, h) | __funnelshift_l(0u, 1u, h >> 5); }
__device__ __forceinline__ bool kh_multi(const uint32_t *multi, uint32_t r) { return (__ldg(multi + (r >> 5)) >> (r & 31)) & 1u; }
__device__ __forceinline__ uint32_t kh_slot(uint32_t x) { return (x * 0x85EBCA6Bu) >> 18; }   // 14 bits
__device__ __forceinline__ uint32_t kh_lds(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 kh_lds2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// number of elements of the sorted array a[0..len) that are < v, searching from `from`
__device__ __forceinline__ uint32_t kh_lower_bound(const uint32_t *a, uint32_t from, uint32_t len, uint32_t v)
{
    uint32_t lo = from, hi = len;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- pivots: one warp per row block; interior pivot q = smallest v with #(elements < v) >= q*total/P
__global__ void k3h_pivot_kernel(const uint32_t *ranks, const unsigned long long *off, const uint32_t *size, uint32_t n,
                                 const uint32_t *pivoff, uint32_t nrb, const uint32_t *distinct, uint32_t *pivots)
{
    const uint32_t rb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (rb >= nrb) return;
    if (distinct[rb * KH_DREC] != 0xFFFFFFFFu) return;      // pivots of this row block came from the sort kernel
    const uint32_t np = pivoff[rb + 1] - pivoff[rb];
    if (np == 0) return;
    const uint32_t id = rb * 32 + lane;
    const uint32_t len = id < n ? size[id] : 0u;
    const uint32_t *a = ranks + (id < n ? off[id] : 0ULL);
    unsigned long long total = len;
    for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    for (uint32_t q = 1; q <= np; q++) {
        const unsigned long long target = total * q / (np + 1);
        uint32_t lo = 0, hi = 0xFFFFFFFEu;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            uint32_t c = kh_lower_bound(a, 0, len, mid);
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (c >= target) hi = mid; else lo = mid + 1;
        }
        if (lane == 0) pivots[pivoff[rb] + q - 1] = lo;
    }
}

// ---- register-blocked passes of the bitonic merges in k3h_distinct_kernel.  A merge of size sz is the strides sz/2 .. 1;
// one pass over shared memory per stride (and a CTA barrier) made the sort 65 passes long.  kd_pass_high does B strides
// 2^(j+B-1) .. 2^j (j >= 5) per pass: a thread holds the 2^B elements that differ in bits [j, j+B) and exchanges them in
// registers (lanes walk consecutive addresses: no bank conflicts).  kd_pass_low does the nine strides 256 .. 1: a warp
// owns a 512-element chunk, lane l holds elements l + 32k (k < 16); strides 256 .. 32 are register exchanges, 16 .. 1
// warp shuffles.  Both need sz >= 512, so the direction bit (idx & sz) is the same for everything a thread touches.
__device__ __forceinline__ void kd_cx(uint32_t &x, uint32_t &y, uint32_t &rx, uint32_t &ry, bool asc)
{
    if ((x > y) == asc) { uint32_t t = x; x = y; y = t; t = rx; rx = ry; ry = t; }
}
template <int B>
__device__ __forceinline__ void kd_pass_high(uint32_t *buf, uint8_t *rowof, uint32_t npad, uint32_t sz, uint32_t j, uint32_t tid)
{
    constexpr int E = 1 << B;
    for (uint32_t t = tid; t < npad / E; t += KH_THREADS) {
        const uint32_t base = ((t >> j) << (j + B)) | (t & ((1u << j) - 1u));
        const bool asc = (base & sz) == 0;
        uint32_t v[E], r[E];
#pragma unroll
        for (int k = 0; k < E; k++) { v[k] = buf[base + ((uint32_t)k << j)]; r[k] = rowof[base + ((uint32_t)k << j)]; }
#pragma unroll
        for (int m = B - 1; m >= 0; m--)
#pragma unroll
            for (int k = 0; k < E; k++)
                if (!(k & (1 << m))) kd_cx(v[k], v[k | (1 << m)], r[k], r[k | (1 << m)], asc);
#pragma unroll
        for (int k = 0; k < E; k++) { buf[base + ((uint32_t)k << j)] = v[k]; rowof[base + ((uint32_t)k << j)] = (uint8_t)r[k]; }
    }
}
__device__ __forceinline__ void kd_pass_low(uint32_t *buf, uint8_t *rowof, uint32_t npad, uint32_t sz, uint32_t warp, uint32_t lane)
{
    for (uint32_t cb = warp * 512u; cb < npad; cb += (KH_THREADS / 32) * 512u) {
        const bool asc = (cb & sz) == 0;
        uint32_t v[16], r[16];
#pragma unroll
        for (int k = 0; k < 16; k++) { v[k] = buf[cb + lane + 32u * k]; r[k] = rowof[cb + lane + 32u * k]; }
#pragma unroll
        for (int m = 3; m >= 0; m--)
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (!(k & (1 << m))) kd_cx(v[k], v[k | (1 << m)], r[k], r[k | (1 << m)], asc);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const bool upper = (lane & s) != 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t ov = __shfl_xor_sync(0xffffffffu, v[k], s), orow = __shfl_xor_sync(0xffffffffu, r[k], s);
                const uint32_t x = upper ? ov : v[k], y = upper ? v[k] : ov;     // x sits at the lower index
                if ((x > y) == asc) { v[k] = ov; r[k] = orow; }
            }
        }
#pragma unroll
        for (int k = 0; k < 16; k++) { buf[cb + lane + 32u * k] = v[k]; rowof[cb + lane + 32u * k] = (uint8_t)r[k]; }
    }
}

// ---- distinct-aware partitions.  The table holds DISTINCT ranks, and the 32 sketches of a row block often share
// most of theirs (neighbouring genomes of one species): 32 x 1000 elements are then ~9000 distinct keys and fit one
// partition instead of three.  One CTA sorts the row block's elements in shared memory (bitonic, <= 32768 u32),
// counts the distinct ones and emits, for P = 2, 3, 4, the keys at distinct rank q*D/P as pivot candidates.
// out[rb*8] = D (0xFFFFFFFF: block too large, the multiset-quantile kernel above takes over), out[rb*8+1..6] =
// candidates for (P,q) = (2,1) (3,1) (3,2) (4,1) (4,2) (4,3), out[rb*8+7] = elements that entered the sort.
__global__ void __launch_bounds__(KH_THREADS, 1) k3h_distinct_kernel(const uint32_t *ranks, const unsigned long long *off,
                                                                      const uint32_t *size, uint32_t n, const uint32_t *multi,
                                                                      uint32_t *out, uint32_t *ukeys, uint32_t *umask, int blocked,
                                                                      const uint32_t *block_list)
{
    extern __shared__ __align__(16) uint8_t kh_smem[];
    uint32_t *buf = reinterpret_cast<uint32_t *>(kh_smem);
    uint8_t *rowof = kh_smem + (size_t)KH_SORT_CAP * 4;          // which of the 32 sketches an element came from
    __shared__ uint32_t s_pos[33], s_warp[32], s_d, s_cnt[32];
    __shared__ unsigned long long s_raw;
    const uint32_t rb = block_list ? block_list[blockIdx.x] : blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        unsigned long long raw = 0;
        for (uint32_t l = 0; l < 32; l++) { const uint32_t id = rb * 32 + l; raw += id < n ? size[id] : 0u; }
        s_raw = raw;
    }
    __syncthreads();
    // blocks of very large sketches are not worth the counting pass: no list, K3H handles them
    if (s_raw > 8ull * KH_SORT_CAP) { if (tid == 0) out[rb * KH_DREC] = 0xFFFFFFFFu; return; }
    // Only ranks that occur in at least two sketches of the whole set can ever be shared (bitmap `multi`, from the rank
    // transform): the others -- strain-private k-mers, a quarter of every sketch in the bench model and far more
    // between unrelated genomes -- take no part in any intersection and are left out of the lists and of the tables.
    const uint32_t my_id = rb * 32 + warp;
    const uint32_t my_len = my_id < n ? size[my_id] : 0u;
    const uint32_t *my_row = ranks + (my_id < n ? off[my_id] : 0ULL);
    {   // warp w counts the shared elements of row w
        uint32_t c = 0;
        for (uint32_t e0 = 0; e0 < my_len; e0 += 32) {
            const uint32_t e = e0 + lane;
            const bool keep = e < my_len && kh_multi(multi, __ldg(my_row + e));
            c += __popc(__ballot_sync(0xffffffffu, keep));
        }
        if (lane == 0) s_cnt[warp] = c;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t l = 0; l < 32; l++) { s_pos[l] = acc; acc += s_cnt[l]; }
        s_pos[32] = acc;
    }
    __syncthreads();
    const uint32_t total = s_pos[32];
    if (tid == 0) out[rb * KH_DREC + 7] = total;                 // shared elements of the block
    if (total > KH_SORT_CAP) { if (tid == 0) out[rb * KH_DREC] = 0xFFFFFFFFu; return; }
    // Every row arrives sorted.  When the longest one fits R <= 1024 slots (R a power of two), row w goes into slots
    // [w*R, (w+1)*R) padded with 0xFFFFFFFF -- ascending for even w, descending for odd w -- which is the state of a
    // bitonic sort after its size-R stage: only the merges of size 2R .. 32R remain (65 of 120 stages at R = 1024).
    uint32_t maxlen = 0;
    for (uint32_t l = 0; l < 32; l++) maxlen = max(maxlen, s_cnt[l]);
    uint32_t R = 32;
    while (R < maxlen) R <<= 1;
    const bool presorted = R * 32 <= KH_SORT_CAP;
    uint32_t npad = 1024, first_sz = 2;
    if (presorted) { npad = R * 32; first_sz = 2 * R; }
    else while (npad < total) npad <<= 1;
    {   // warp w compacts the shared elements of row w into its slots
        const uint32_t cnt = s_cnt[warp];
        const uint32_t base = presorted ? warp * R : s_pos[warp];
        const bool reversed = presorted && (warp & 1);
        uint32_t done = 0;
        for (uint32_t e0 = 0; e0 < my_len; e0 += 32) {
            const uint32_t e = e0 + lane;
            const uint32_t x = e < my_len ? __ldg(my_row + e) : 0u;
            const bool keep = e < my_len && kh_multi(multi, x);
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const uint32_t idx = done + __popc(b & ((1u << lane) - 1));
                const uint32_t slot = base + (reversed ? R - 1 - idx : idx);
                buf[slot] = x; rowof[slot] = (uint8_t)warp;
            }
            done += __popc(b);
        }
        if (presorted) {
            for (uint32_t e = cnt + lane; e < R; e += 32) {
                const uint32_t slot = base + (reversed ? R - 1 - e : e);
                buf[slot] = 0xFFFFFFFFu; rowof[slot] = 0;
            }
        } else {
            for (uint32_t e = total + tid; e < npad; e += KH_THREADS) { buf[e] = 0xFFFFFFFFu; rowof[e] = 0; }
        }
    }
    __syncthreads();
    for (uint32_t sz = first_sz; sz <= npad; sz <<= 1) {
        if (blocked && sz >= 512) {
            // strides sz/2 .. 512 in register-blocked passes of up to four strides, then 256 .. 1 in one pass
            int top = 31 - __clz(sz) - 1;                        // exponent of the largest stride
            while (top >= 9) {
                const int B = min(4, top - 8);
                const uint32_t j = (uint32_t)(top - B + 1);
                if (B == 4) kd_pass_high<4>(buf, rowof, npad, sz, j, tid);
                else if (B == 3) kd_pass_high<3>(buf, rowof, npad, sz, j, tid);
                else if (B == 2) kd_pass_high<2>(buf, rowof, npad, sz, j, tid);
                else kd_pass_high<1>(buf, rowof, npad, sz, j, tid);
                __syncthreads();
                top -= B;
            }
            kd_pass_low(buf, rowof, npad, sz, warp, lane);
            __syncthreads();
            continue;
        }
        for (uint32_t stride = sz >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (npad >> 1); t += KH_THREADS) {
                const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const uint32_t x = buf[lo], y = buf[hi];
                if ((x > y) == ((lo & sz) == 0)) {
                    buf[lo] = y; buf[hi] = x;
                    const uint8_t rx = rowof[lo]; rowof[lo] = rowof[hi]; rowof[hi] = rx;
                }
            }
            __syncthreads();
        }
    }
    // distinct count: thread t owns the chunk [t*c, (t+1)*c)
    // (c is a power of two <= 32; the chunk is walked from a lane-dependent start so that the 32 lanes of a warp read 32
    // banks -- in chunk order every lane would hit the same bank: 43 % of the kernel's shared-memory wavefronts)
    const uint32_t c = npad / KH_THREADS, b0 = tid * c;
    uint32_t heads = 0;                                           // bit k: element b0 + k starts a run of equal keys
    for (uint32_t kk = 0; kk < c; kk++) {
        const uint32_t k = (kk + lane) & (c - 1), i = b0 + k;
        if (i < total && (i == 0 || buf[i] != buf[i - 1])) heads |= 1u << k;
    }
    const uint32_t mine = __popc(heads);
    uint32_t incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += u; }
        s_warp[lane] = v;
        if (lane == 31) s_d = v;
    }
    __syncthreads();
    const uint32_t D = s_d;
    const uint32_t before = incl - mine + (warp ? s_warp[warp - 1] : 0u); // distinct keys before this chunk
    if (tid == 0) out[rb * KH_DREC] = D;
    // wanted distinct ranks
    uint32_t want[6];
    want[0] = (uint32_t)((unsigned long long)D * 1 / 2);
    want[1] = (uint32_t)((unsigned long long)D * 1 / 3); want[2] = (uint32_t)((unsigned long long)D * 2 / 3);
    want[3] = (uint32_t)((unsigned long long)D * 1 / 4); want[4] = (uint32_t)((unsigned long long)D * 2 / 4);
    want[5] = (uint32_t)((unsigned long long)D * 3 / 4);
    // union list of the block (K3J): distinct key d -> (key, mask of the sketches that hold it), ascending, stored at
    // the element offset of the block's first sketch (a block never has more distinct keys than elements)
    const unsigned long long base = rb * 32 < n ? off[rb * 32] : 0ULL;
    for (uint32_t kk = 0; kk < c; kk++) {
        const uint32_t k = (kk + lane) & (c - 1);
        if (!((heads >> k) & 1u)) continue;
        const uint32_t i = b0 + k, d = before + __popc(heads & ((1u << k) - 1u));   // distinct rank of this key
        const uint32_t key = buf[i];
#pragma unroll
        for (int q = 0; q < 6; q++) if (d == want[q]) out[rb * KH_DREC + 1 + q] = key;
        if (ukeys) {
            uint32_t m = 0;
            for (uint32_t j = i; j < total && buf[j] == key; j++) m |= 1u << rowof[j];
            ukeys[base + d] = key;
            umask[base + d] = m;
        }
    }
}

// ---- union lists through a hash table (round 2; the sort above is the fall-back).  The 32 sketches of a block of related
// genomes hold ~24,000 shared elements but only ~1,300 DISTINCT ranks: instead of sorting all elements (65 bitonic stages
// over 32,768 slots) every element is inserted into a shared-memory table rank -> mask of the rows holding it (the very
// table K3H/K3J build), the occupied slots are compacted and only the distinct keys are sorted (66 stages over 2,048
// slots).  Blocks with more than KDH_CAP distinct shared ranks answer 0xFFFFFFFE and are redone by k3h_distinct_kernel.
// Same outputs: out[rb*8] = D, [1..6] pivot candidates, [7] = shared elements; (key, mask) lists at off[rb*32].
constexpr uint32_t KDH_SLOTS = 16384;
constexpr uint32_t KDH_CAP = 8192;
constexpr size_t KDH_SMEM = (size_t)KDH_SLOTS * 8 + (size_t)KDH_CAP * 8;
constexpr uint32_t KDH_RETRY = 0xFFFFFFFEu;

__global__ void __launch_bounds__(KH_THREADS, 1) k3h_distinct_hash_kernel(const uint32_t *ranks, const unsigned long long *off,
                                                                           const uint32_t *size, uint32_t n, const uint32_t *multi,
                                                                           uint32_t *out, uint32_t *ukeys, uint32_t *umask)
{
    extern __shared__ __align__(16) uint8_t kh_smem[];
    uint2 *table = reinterpret_cast<uint2 *>(kh_smem);
    unsigned long long *sk = reinterpret_cast<unsigned long long *>(kh_smem + (size_t)KDH_SLOTS * 8);
    __shared__ uint32_t s_elems, s_distinct, s_over, s_warp[32];
    const uint32_t rb = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        uint4 *t4 = reinterpret_cast<uint4 *>(kh_smem);
        const uint4 e = make_uint4(KH_EMPTY, 0u, KH_EMPTY, 0u);
        for (uint32_t i = tid; i < KDH_SLOTS / 2; i += KH_THREADS) t4[i] = e;
        if (tid == 0) { s_elems = 0; s_distinct = 0; s_over = 0; }
    }
    __syncthreads();
    {   // warp w inserts the shared elements of row w
        const uint32_t id = rb * 32 + warp;
        const uint32_t len = id < n ? size[id] : 0u;
        const uint32_t *row = ranks + (id < n ? off[id] : 0ULL);
        const uint32_t bit = 1u << warp;
        uint32_t mine = 0;
        for (uint32_t e0 = 0; e0 < len; e0 += 128) {
            uint32_t x[4]; bool keep[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {                          // four independent loads and bitmap look-ups in flight
                const uint32_t e = e0 + 32 * q + lane;
                x[q] = e < len ? __ldg(row + e) : 0u;
                keep[q] = e < len;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) keep[q] = keep[q] && kh_multi(multi, x[q]);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (!keep[q] || *(volatile uint32_t *)&s_over) continue;
                mine++;
                uint32_t slot = kh_slot(x[q]);
                for (;;) {
                    // most elements find their key in place already (24 of the 32 rows hold it): a plain read first, the
                    // compare-and-swap only for an empty slot
                    uint32_t old = *(volatile uint32_t *)&table[slot].x;
                    if (old == KH_EMPTY) {
                        old = atomicCAS(&table[slot].x, KH_EMPTY, x[q]);
                        if (old == KH_EMPTY && atomicAdd(&s_distinct, 1u) + 1u > KDH_CAP) s_over = 1u;
                        if (old == KH_EMPTY) old = x[q];
                    }
                    if (old == x[q]) { atomicOr(&table[slot].y, bit); break; }
                    slot = (slot + 1) & (KDH_SLOTS - 1);
                }
            }
        }
        for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0 && mine) atomicAdd(&s_elems, mine);
    }
    __syncthreads();
    if (s_over) { if (tid == 0) out[rb * KH_DREC] = KDH_RETRY; return; }
    const uint32_t D = s_distinct;
    // compact the occupied slots: thread t owns slots t, t + 1024, ... (consecutive lanes read consecutive slots: no bank
    // conflicts; the order of the compacted entries does not matter, they are sorted next)
    uint32_t cnt = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) cnt += table[q * KH_THREADS + tid].x != KH_EMPTY;
    uint32_t incl = cnt;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += u; }
        s_warp[lane] = v;
    }
    __syncthreads();
    uint32_t pos = incl - cnt + (warp ? s_warp[warp - 1] : 0u);
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const uint2 e = table[q * KH_THREADS + tid];
        if (e.x != KH_EMPTY) sk[pos++] = ((unsigned long long)e.x << 32) | e.y;
    }
    uint32_t npad = 32;
    while (npad < D) npad <<= 1;
    for (uint32_t i = D + tid; i < npad; i += KH_THREADS) sk[i] = 0xFFFFFFFFFFFFFFFFull;
    __syncthreads();
    // bitonic sort of the distinct keys (key in the high word)
    for (uint32_t sz = 2; sz <= npad; sz <<= 1)
        for (uint32_t stride = sz >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (npad >> 1); t += KH_THREADS) {
                const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const unsigned long long x = sk[lo], y = sk[hi];
                if ((x > y) == ((lo & sz) == 0)) { sk[lo] = y; sk[hi] = x; }
            }
            __syncthreads();
        }
    const unsigned long long base = rb * 32 < n ? off[rb * 32] : 0ULL;
    for (uint32_t d = tid; d < D; d += KH_THREADS) {
        const unsigned long long v = sk[d];
        ukeys[base + d] = (uint32_t)(v >> 32);
        umask[base + d] = (uint32_t)v;
    }
    if (tid < 8) {
        uint32_t v;
        if (tid == 0) v = D;
        else if (tid == 7) v = s_elems;
        else {
            const uint32_t num = tid == 1 ? 1 : tid == 2 ? 1 : tid == 3 ? 2 : tid == 4 ? 1 : tid == 5 ? 2 : 3;
            const uint32_t den = tid == 1 ? 2 : tid <= 3 ? 3 : 4;
            const uint32_t w = (uint32_t)((unsigned long long)D * num / den);
            v = w < D ? (uint32_t)(sk[w] >> 32) : 0u;
        }
        out[rb * KH_DREC + tid] = v;
    }
}

template <int MODE>
__global__ void __launch_bounds__(KH_THREADS, 1) k3h_join_kernel(const KHArgs a)
{
    constexpr bool CAND = MODE == KH_MODE_CAND;
    constexpr uint32_t STRIDE = CAND ? KH_STATE_CAND : KH_STATE;
    extern __shared__ __align__(16) uint8_t kh_smem[];
    uint2 *table = reinterpret_cast<uint2 *>(kh_smem);
    uint32_t *filter = reinterpret_cast<uint32_t *>(kh_smem + (size_t)KH_SLOTS * 8);
    __shared__ uint32_t s_unit, s_next;
    __shared__ uint32_t s_lo[32], s_hi[32], s_len[32];
    __shared__ unsigned long long s_base[32];

    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t lane = tid & 31;
    asm volatile("" : "+r"(lane));
    uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(table);
    asm volatile("" : "+r"(tab_addr));
    uint32_t filt_addr = (uint32_t)__cvta_generic_to_shared(filter);
    asm volatile("" : "+r"(filt_addr));      // keep the shared-window addresses in registers (no rematerialisation)
    uint32_t *state = a.state + (size_t)blockIdx.x * KH_CSEG * STRIDE;

    for (;;) {
        if (tid == 0) s_unit = (uint32_t)(atomicAdd(&a.counters[2], 1ULL) * a.nparts + a.part);
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= a.nunits) break;
        const tc_kh_unit unit = a.units[u];
        const uint32_t rb = unit.rb, c0 = unit.c0, c1 = unit.c1;
        const uint32_t piv0 = a.pivoff[rb];
        const uint32_t P = a.pivoff[rb + 1] - piv0 + 1;
        if (tid < 32) {
            const uint32_t id = rb * 32 + tid;
            s_len[tid] = id < a.n ? a.size[id] : 0u;
            s_base[tid] = id < a.n ? a.off[id] : 0ULL;
            s_lo[tid] = 0;
        }
        __syncthreads();

        for (uint32_t p = 0; p < P; p++) {
            const uint32_t piv_hi = (p + 1 == P) ? 0xFFFFFFFFu : a.pivots[piv0 + p];
            // ---- clear table + filter, find this partition's end in every row
            {
                uint4 *t4 = reinterpret_cast<uint4 *>(kh_smem);
                const uint4 e = make_uint4(KH_EMPTY, 0u, KH_EMPTY, 0u);
                for (uint32_t i = tid; i < KH_SLOTS / 2; i += KH_THREADS) t4[i] = e;
                for (uint32_t i = tid; i < KH_FWORDS; i += KH_THREADS) filter[i] = 0xFFFFFFFFu;   // stored inverted: 0 = present
                if (tid < 32) {
                    const uint32_t *row = a.ranks + s_base[tid];
                    s_hi[tid] = (p + 1 == P) ? s_len[tid] : kh_lower_bound(row, s_lo[tid], s_len[tid], piv_hi);
                }
                if (tid == 0) s_next = c0;
            }
            __syncthreads();
            // ---- build: warp w inserts the partition's elements of row w
            {
                const uint32_t *row = a.ranks + s_base[warp];
                const uint32_t bit = 1u << warp;
                for (uint32_t e = s_lo[warp] + lane; e < s_hi[warp]; e += 32) {
                    const uint32_t x = __ldg(row + e);
                    if (!kh_multi(a.multi, x)) continue;          // held by this sketch only: cannot be shared
                    const uint32_t h = x * 0x9E3779B1u;
                    atomicAnd(&filter[kh_fword(h)], ~kh_fmask(h));
                    uint32_t slot = kh_slot(x);
                    for (;;) {
                        const uint32_t old = atomicCAS(&table[slot].x, KH_EMPTY, x);
                        if (old == KH_EMPTY || old == x) { atomicOr(&table[slot].y, bit); break; }
                        slot = (slot + 1) & (KH_SLOTS - 1);
                    }
                }
            }
            __syncthreads();
            // ---- stream the columns of the segment through the table, one warp per column, 128 elements
            //      (4 coalesced 128-byte loads) per step; the next step's loads are issued before the probes
            for (;;) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(&s_next, 1u);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= c1) break;
                if (!CAND && a.tile_flags) {           // no rank of this column's block occurs in the row block: every count is 0
                    const uint32_t cb = c >> 5;
                    if (!((__ldg(a.tile_flags + (size_t)rb * a.tile_words + (cb >> 5)) >> (cb & 31)) & 1u)) continue;
                }
                const uint32_t cid = (CAND && a.col_ids) ? __ldg(a.col_ids + c) : c;      // sketch streamed as column c
                const uint32_t clen = a.size[cid];
                uint32_t *st = state + (size_t)(c - c0) * STRIDE;
                uint32_t pos = 0, cnt = 0;
                uint32_t first_x = 0xFFFFFFFFu;        // CAND: smallest shared rank with row `lane` (= first one met)
                if (p > 0) { pos = st[32]; cnt = st[lane]; if (CAND) first_x = st[33 + lane]; }
                const uint32_t *pc = a.ranks + a.off[cid] + pos + lane;   // this lane's element of the current step
                uint32_t rem = clen - pos;                               // elements left from the step's start
                uint32_t xa[4], xb[4];
#pragma unroll
                for (int k = 0; k < 4; k++) xa[k] = (lane + 32 * k < rem) ? __ldg(pc + 32 * k) : 0xFFFFFFFFu;
                // one step = the 128 elements in x; xn receives the next step's elements (loads issued before the
                // probes).  Returns false when the partition ends inside this step.
                auto step = [&](uint32_t (&x)[4], uint32_t (&xn)[4]) -> bool {
                    if (rem >= 256u) {                       // warp-uniform: the whole next step is in bounds
#pragma unroll
                        for (int k = 0; k < 4; k++) xn[k] = __ldg(pc + 128 + 32 * k);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) xn[k] = (lane + 128 + 32 * k < rem) ? __ldg(pc + 128 + 32 * k) : 0xFFFFFFFFu;
                    }
                    // presence filter (blocked Bloom, 2 bits in one word): 4 independent LDS per lane
                    uint32_t fw[4], fm[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t h = x[k] * 0x9E3779B1u;
                        fw[k] = kh_lds(filt_addr + ((h >> (30 - KH_FBITS + 5)) & ((KH_FWORDS - 1) << 2)));
                        fm[k] = kh_fmask(h);
                    }
                    bool maybe = false;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        // no range test: an element >= piv_hi (or a tail sentinel) that slips through the filter
                        // simply misses in the table, which only holds keys of this partition
                        fw[k] = (fm[k] & fw[k]) == 0u;          // the filter words are stored inverted
                        maybe |= fw[k] != 0u;
                    }
                    uint32_t mask[4] = {0u, 0u, 0u, 0u};
                    if (maybe) {                               // table lookups for the filter hits
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            if (fw[k]) {
                                uint32_t slot = kh_slot(x[k]);
                                for (;;) {
                                    const uint2 e = kh_lds2(tab_addr + slot * 8);
                                    if (e.x == x[k]) { mask[k] = e.y; break; }
                                    if (e.x == KH_EMPTY) break;
                                    slot = (slot + 1) & (KH_SLOTS - 1);
                                }
                            }
                        }
                    }
                    if (__ballot_sync(0xffffffffu, (mask[0] | mask[1] | mask[2] | mask[3]) != 0u)) {
#pragma unroll 1
                        for (int k = 0; k < 4; k++) {
                            const uint32_t mk = k == 0 ? mask[0] : k == 1 ? mask[1] : k == 2 ? mask[2] : mask[3];
                            const uint32_t xk = k == 0 ? x[0] : k == 1 ? x[1] : k == 2 ? x[2] : x[3];
                            unsigned hit = __ballot_sync(0xffffffffu, mk != 0u);
                            if (__popc(hit) <= 16) {
                                while (hit) {
                                    const int src = __ffs(hit) - 1;
                                    const uint32_t m = __shfl_sync(0xffffffffu, mk, src);
                                    cnt += (m >> lane) & 1u;
                                    if (CAND) {        // elements arrive in ascending order: the first hit is the smallest
                                        const uint32_t xv = __shfl_sync(0xffffffffu, xk, src);
                                        if (((m >> lane) & 1u) && first_x == 0xFFFFFFFFu) first_x = xv;
                                    }
                                    hit &= hit - 1;
                                }
                            } else {
#pragma unroll
                                for (int r = 0; r < 32; r++) {
                                    const unsigned bb = __ballot_sync(0xffffffffu, (mk >> r) & 1u);
                                    if (lane == r) cnt += __popc(bb);
                                    if (CAND) {
                                        const uint32_t xv = __shfl_sync(0xffffffffu, xk, bb ? __ffs(bb) - 1 : 0);
                                        if (lane == r && bb && first_x == 0xFFFFFFFFu) first_x = xv;
                                    }
                                }
                            }
                        }
                    }
                    // sorted stream: the step is consumed completely iff its last element is in the partition
                    if (__shfl_sync(0xffffffffu, x[3], 31) < piv_hi) {
                        pos += 128; pc += 128; rem -= 128;
                        return true;
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) pos += __popc(__ballot_sync(0xffffffffu, x[k] < piv_hi));
                    return false;
                };
                // (instantiating the step twice with the register sets swapped saves the four moves but doubles the
                // loop body; measured 3 % slower)
                for (;;) {
                    if (!step(xa, xb)) break;
#pragma unroll
                    for (int k = 0; k < 4; k++) xa[k] = xb[k];
                }
                if (p + 1 < P) {
                    st[lane] = cnt;
                    if (CAND) st[33 + lane] = first_x;
                    if (lane == 0) st[32] = pos;
                } else {
                    // ---- epilogue: lane r holds |row r n column c|
                    const uint32_t i = rb * 32 + lane, j = cid;
                    if (CAND) {
                        bool keep = false;
                        if (i >= a.q0 && i < a.q1 && cnt >= 1u && (!a.lower_only || j < i)) {
                            const uint32_t s0 = s_len[lane], s1 = clen;
                            int cmin = a.cmin_fixed;
                            if (cmin < 0) {   // greedy.cpp:1211-1219, IEEE-exact operations in the host's order
                                const double t = a.containment
                                    ? __dmul_rn(a.jmin, (double)min(s0, s1))
                                    : __ddiv_rn(__dmul_rn(a.jmin, (double)(s0 + s1)), a.one_plus_jmin);
                                cmin = (int)ceil(t);
                            }
                            keep = (long long)cnt + 1 >= (long long)cmin;
                        }
                        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                        if (ballot) {
                            unsigned long long base = 0;
                            if (lane == 0) base = atomicAdd(&a.counters[0], (unsigned long long)__popc(ballot));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if (keep) {
                                const unsigned long long idx = base + __popc(ballot & ((1u << lane) - 1));
                                if (idx < a.edge_cap) {
                                    tc_cand cd; cd.query = i; cd.rep = j; cd.common = cnt; cd.first_hit = first_x;
                                    a.cands[idx] = cd;
                                } else {
                                    a.counters[1] = 1ULL;
                                }
                            }
                        }
                    } else if (a.mode == KH_MODE_MATRIX) {
                        if (i >= a.i0 && i < a.i1 && j >= a.j0 && j < a.j1)
                            a.out[(size_t)(i - a.i0) * (a.j1 - a.j0) + (j - a.j0)] = i == j ? s_len[lane] : cnt;   // ranks private to one sketch are not in the table
                    } else {
                        bool keep = false;
                        if (i < a.n && i >= a.row_min && j < i && cnt >= 1u) {
                            const uint32_t s0 = s_len[lane], s1 = clen;
                            const uint32_t mn = min(s0, s1), mx = max(s0, s1);
                            // reference: skip if max_size > radio * min_size (int arithmetic), MST.cpp:1481-1487
                            keep = mn > 0 && !((long long)mx > (long long)a.radio * (long long)mn);
                        }
                        if (a.uf && keep) { kh_uf_edge(a, i, j, cnt, s_len[lane], clen); keep = false; }
                        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                        if (ballot) {
                            unsigned long long base = 0;
                            if (lane == 0) base = atomicAdd(&a.counters[0], (unsigned long long)__popc(ballot));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if (keep) {
                                const unsigned long long idx = base + __popc(ballot & ((1u << lane) - 1));
                                if (idx < a.edge_cap) {
                                    tc_pair e; e.i = i; e.j = j; e.common = cnt;
                                    a.edges[idx] = e;
                                } else {
                                    a.counters[1] = 1ULL;
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (tid < 32) s_lo[tid] = s_hi[tid];
        }
        __syncthreads();
    }
}

}  // namespace

// =====================================================================================================
// K3J  k3j_block_kernel -- the dense pass as a join of 32-sketch BLOCKS on both sides.
//
// K3H streams every column sketch through the row block's table on its own: each (column, partition) visit has a
// fixed cost, a partition slice of one sketch is only ~2.5 steps long, and 32 related columns probe the same keys 32
// times.  Here the column side is a block as well: k3h_distinct_kernel leaves, per block of 32 sketches, the sorted
// list of its DISTINCT ranks with a 32-bit mask of the member sketches that hold each.  A warp takes a (row block x
// column block) tile, streams that one list (9,000 keys for 32 related sketches, 32,000 for unrelated ones, instead of
// 32 x 1,000) through the table, and on a hit adds 1 to count[r][c] for every r in the row mask and c in the column
// mask.  The 32 x 32 counters of the tile are u16 in shared memory (2 KB per warp); hits are rare outside related
// tiles, and a tile without any hit skips its epilogue.  Partition slices of a sorted list are index ranges, so there
// is no per-column stream state; with P > 1 partitions the tile's counters are parked in the per-CTA scratch.
// Column blocks too large for a union list (bu_count = 0xFFFFFFFF) are routed to K3H by the host.
constexpr uint32_t KJ_FBITS = 18;                                   // 32 KB presence filter (64 KB go to the counters)
constexpr uint32_t KJ_FWORDS = (1u << KJ_FBITS) / 32;
constexpr uint32_t KJ_TILE_WORDS = 512;                             // 32 x 32 u16 counters
constexpr uint32_t KJ_TILE_STATE = KJ_TILE_WORDS + 1;               // + "tile had a hit" flag
constexpr size_t KJ_SMEM = (size_t)KH_SLOTS * 8 + (size_t)KJ_FWORDS * 4 + (size_t)32 * KJ_TILE_WORDS * 4;
static_assert(KJ_SMEM <= 227 * 1024 - 1024, "K3J shared memory");
// K3J units are wider than K3H's: the per-unit costs (table build, the wait for the last warp at the unit barrier) are
// a quarter of the pass at 4096 columns per unit (59.4 ms -> 51.1 ms at 8192)
// (46.8 ms at 16384).  The width is chosen per pass (kj_pick_cseg): as wide as possible while every CTA still gets
// enough units for the dynamic hand-out to even out the tail.
constexpr uint32_t KJ_CSEG_MAX = 32768;                                  // widest unit of a partitioned row block (P > 1): what the per-CTA tile state holds
constexpr uint32_t KJ_CSEG_WIDEST = 1u << 20;                            // widest unit kj_pick_cseg chooses (TCLUST_KJ_CSEG_MAX)
constexpr uint32_t KJ_HEAVY_CAP = 256;                                  // deferred tiles per unit and partition
constexpr size_t KJ_STATE_WORDS = (size_t)(KJ_CSEG_MAX / 32) * KJ_TILE_STATE;   // per-CTA tile state (P > 1 only)
constexpr size_t KH_STATE_WORDS_MAX = KJ_STATE_WORDS > (size_t)KH_CSEG * KH_STATE_CAND ? KJ_STATE_WORDS : (size_t)KH_CSEG * KH_STATE_CAND;

__device__ __forceinline__ uint32_t kj_fword(uint32_t h) { return h >> (32 - (KJ_FBITS - 5)); }

__global__ void __launch_bounds__(KH_THREADS, 1) k3j_block_kernel(const KHArgs a)
{
    extern __shared__ __align__(16) uint8_t kh_smem[];
    uint2 *table = reinterpret_cast<uint2 *>(kh_smem);
    uint32_t *filter = reinterpret_cast<uint32_t *>(kh_smem + (size_t)KH_SLOTS * 8);
    uint32_t *cnt_all = filter + KJ_FWORDS;
    __shared__ uint32_t s_unit, s_next, s_nheavy, s_heavy[KJ_HEAVY_CAP];
    __shared__ uint32_t s_lo[32], s_hi[32], s_len[32];
    __shared__ unsigned long long s_base[32];
    __shared__ uint32_t s_rlo, s_rhi, s_nflag;

    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t lane = tid & 31;
    asm volatile("" : "+r"(lane));
    uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(table);
    asm volatile("" : "+r"(tab_addr));
    uint32_t filt_addr = (uint32_t)__cvta_generic_to_shared(filter);
    asm volatile("" : "+r"(filt_addr));
    uint32_t *cntw = cnt_all + warp * KJ_TILE_WORDS;                  // this warp's tile counters, as 512 words
    uint16_t *cnt = reinterpret_cast<uint16_t *>(cntw);              // cnt[r * 32 + c]
    uint32_t *state = a.state + (size_t)blockIdx.x * KH_STATE_WORDS_MAX;

    for (;;) {
        if (tid == 0) s_unit = (uint32_t)(atomicAdd(&a.counters[2], 1ULL) * a.nparts + a.part);
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= a.nunits) break;
        const tc_kh_unit unit = a.units[u];
        const uint32_t rb = unit.rb, c0 = unit.c0, c1 = unit.c1;      // c0 is a multiple of 32
        const uint32_t piv0 = a.pivoff[rb];
        const uint32_t P = a.pivoff[rb + 1] - piv0 + 1;
        const uint32_t cb_begin = c0 / 32, cb_end = (c1 + 31) / 32;
        const uint32_t *flag_row = a.tile_flags ? a.tile_flags + (size_t)rb * a.tile_words : nullptr;
        // the row block's own union list (ascending distinct shared ranks + the mask of the rows holding each): the
        // table is built from it -- ~1,300 insertions for 32 related sketches instead of 32 x 1000 elements with a
        // bitmap look-up each.  0xFFFFFFFF: the block has no list, build from the rows.
        const uint32_t Drow = a.bu_count[rb * KH_DREC];
        const unsigned long long rbase = rb * 32 < a.n ? a.off[rb * 32] : 0ULL;
        const uint32_t *rkeys = a.bu_keys + rbase, *rmasks = a.bu_mask + rbase;
        if (tid < 32) {
            const uint32_t id = rb * 32 + tid;
            s_len[tid] = id < a.n ? a.size[id] : 0u;
            s_base[tid] = id < a.n ? a.off[id] : 0ULL;
            s_lo[tid] = 0;
            // flagged tiles of this unit: with many of them every warp keeps its tiles to itself, with a few the
            // dense ones are shared by the whole CTA (deferred tiles below)
            uint32_t nf = 0;
            if (flag_row) {
                for (uint32_t cb = cb_begin + tid; cb < cb_end; cb += 32) nf += (__ldg(flag_row + (cb >> 5)) >> (cb & 31)) & 1u;
                for (int o = 16; o; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o);
            } else nf = cb_end - cb_begin;
            if (tid == 0) s_nflag = nf;
        }
        __syncthreads();
        const bool unit_defers = s_nflag < 64u;

        for (uint32_t p = 0; p < P; p++) {
            const uint32_t piv_lo = p == 0 ? 0u : a.pivots[piv0 + p - 1];
            const uint32_t piv_hi = (p + 1 == P) ? 0xFFFFFFFFu : a.pivots[piv0 + p];
            {   // clear table + filter, find this partition's end in every row
                uint4 *t4 = reinterpret_cast<uint4 *>(kh_smem);
                const uint4 e = make_uint4(KH_EMPTY, 0u, KH_EMPTY, 0u);
                for (uint32_t i = tid; i < KH_SLOTS / 2; i += KH_THREADS) t4[i] = e;
                for (uint32_t i = tid; i < KJ_FWORDS; i += KH_THREADS) filter[i] = 0xFFFFFFFFu;   // inverted: 0 = present
                if (Drow == 0xFFFFFFFFu) {
                    if (tid < 32) {
                        const uint32_t *row = a.ranks + s_base[tid];
                        s_hi[tid] = (p + 1 == P) ? s_len[tid] : kh_lower_bound(row, s_lo[tid], s_len[tid], piv_hi);
                    }
                } else if (tid == 32) {
                    const uint32_t lo = p == 0 ? 0u : kh_lower_bound(rkeys, 0, Drow, piv_lo);
                    s_rlo = lo; s_rhi = (p + 1 == P) ? Drow : kh_lower_bound(rkeys, lo, Drow, piv_hi);
                }
                if (tid == 0) { s_next = 0; s_nheavy = 0; }
            }
            __syncthreads();
            if (Drow != 0xFFFFFFFFu) {   // build from the union list: distinct keys, the mask is ready
                for (uint32_t e = s_rlo + tid; e < s_rhi; e += KH_THREADS) {
                    const uint32_t x = __ldg(rkeys + e), m = __ldg(rmasks + e);
                    const uint32_t h = x * 0x9E3779B1u;
                    atomicAnd(&filter[kj_fword(h)], ~kh_fmask(h));
                    uint32_t slot = kh_slot(x);
                    for (;;) {
                        if (atomicCAS(&table[slot].x, KH_EMPTY, x) == KH_EMPTY) { table[slot].y = m; break; }
                        slot = (slot + 1) & (KH_SLOTS - 1);
                    }
                }
            } else {   // build: warp w inserts the partition's elements of row w
                const uint32_t *row = a.ranks + s_base[warp];
                const uint32_t bit = 1u << warp;
                for (uint32_t e = s_lo[warp] + lane; e < s_hi[warp]; e += 32) {
                    const uint32_t x = __ldg(row + e);
                    if (!kh_multi(a.multi, x)) continue;          // held by this sketch only: cannot be shared
                    const uint32_t h = x * 0x9E3779B1u;
                    atomicAnd(&filter[kj_fword(h)], ~kh_fmask(h));
                    uint32_t slot = kh_slot(x);
                    for (;;) {
                        const uint32_t old = atomicCAS(&table[slot].x, KH_EMPTY, x);
                        if (old == KH_EMPTY || old == x) { atomicOr(&table[slot].y, bit); break; }
                        slot = (slot + 1) & (KH_SLOTS - 1);
                    }
                }
            }
            __syncthreads();
            // ---- tiles: one warp per column block.  stream_slice runs the keys [lo, hi) of one union list through the
            //      table and adds the hits to this warp's counters; with may_defer it gives up (returns true) at the
            //      first group of 32 keys with more than two hits: such a tile -- related sketches on both sides, a
            //      thousand hits with 32 x 32 increments each -- would keep one warp busy ~50 times longer than the
            //      others wait at the unit barrier, so it is redone afterwards by all 32 warps together.
            auto stream_slice = [&](const uint32_t *keys, const uint32_t *cmasks, uint32_t lo, uint32_t hi, bool may_defer,
                                    uint32_t &anyhit) -> bool {
                if (hi <= lo) return false;
                uint32_t pos = lo, rem = hi - lo;
                const uint32_t *pc = keys + lo + lane;
                uint32_t x[4], xn[4];
#pragma unroll
                for (int k = 0; k < 4; k++) x[k] = (lane + 32 * k < rem) ? __ldg(pc + 32 * k) : 0xFFFFFFFFu;
                for (;;) {
                    if (rem >= 256u) {
#pragma unroll
                        for (int k = 0; k < 4; k++) xn[k] = __ldg(pc + 128 + 32 * k);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) xn[k] = (lane + 128 + 32 * k < rem) ? __ldg(pc + 128 + 32 * k) : 0xFFFFFFFFu;
                    }
                    uint32_t fw[4], fm[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t h = x[k] * 0x9E3779B1u;
                        fw[k] = kh_lds(filt_addr + ((h >> (30 - KJ_FBITS + 5)) & ((KJ_FWORDS - 1) << 2)));
                        fm[k] = kh_fmask(h);
                    }
                    bool maybe = false;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        fw[k] = (fm[k] & fw[k]) == 0u;          // filter words are stored inverted
                        maybe |= fw[k] != 0u;
                    }
                    uint32_t mask[4] = {0u, 0u, 0u, 0u};
                    if (maybe) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            if (fw[k]) {
                                uint32_t slot = kh_slot(x[k]);
                                for (;;) {
                                    const uint2 e = kh_lds2(tab_addr + slot * 8);
                                    if (e.x == x[k]) { mask[k] = e.y; break; }
                                    if (e.x == KH_EMPTY) break;
                                    slot = (slot + 1) & (KH_SLOTS - 1);
                                }
                            }
                        }
                    }
                    if (__ballot_sync(0xffffffffu, (mask[0] | mask[1] | mask[2] | mask[3]) != 0u)) {
                        anyhit = 1u;
                        // column masks of the hit keys: every hit lane loads its own (one latency per step)
                        uint32_t cmv[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) cmv[k] = mask[k] ? __ldg(cmasks + pos + lane + 32 * k) : 0u;
#pragma unroll 1
                        for (int k = 0; k < 4; k++) {
                            const uint32_t mk = k == 0 ? mask[0] : k == 1 ? mask[1] : k == 2 ? mask[2] : mask[3];
                            const uint32_t ck = k == 0 ? cmv[0] : k == 1 ? cmv[1] : k == 2 ? cmv[2] : cmv[3];
                            unsigned hit = __ballot_sync(0xffffffffu, mk != 0u);
                            if (__popc(hit) <= 2) {
                                while (hit) {                         // few hits: count[r][c]++ for r in rows, c = lane
                                    const int src = __ffs(hit) - 1;
                                    uint32_t rm = __shfl_sync(0xffffffffu, mk, src);
                                    const uint32_t cm = __shfl_sync(0xffffffffu, ck, src);
                                    const bool mine = (cm >> lane) & 1u;
                                    while (rm) {
                                        const int r = __ffs(rm) - 1;
                                        if (mine) cnt[r * 32 + lane]++;
                                        rm &= rm - 1;
                                    }
                                    hit &= hit - 1;
                                }
                            } else {
                                if (may_defer) return true;
                                // many hits: transpose the 32 keys' masks with ballots -- mycols: keys held by column
                                // `lane`; per row r the keys held by row r -- and add popcounts
                                uint32_t mycols = 0;
#pragma unroll
                                for (int c = 0; c < 32; c++) {
                                    const unsigned bb = __ballot_sync(0xffffffffu, (ck >> c) & 1u);
                                    if (lane == (uint32_t)c) mycols = bb;
                                }
                                uint32_t rows_any = mk;
#pragma unroll
                                for (int o = 16; o; o >>= 1) rows_any |= __shfl_xor_sync(0xffffffffu, rows_any, o);
                                while (rows_any) {
                                    const int r = __ffs(rows_any) - 1;
                                    const unsigned rk = __ballot_sync(0xffffffffu, (mk >> r) & 1u);
                                    const uint32_t add = __popc(rk & mycols);
                                    if (add) cnt[r * 32 + lane] += (uint16_t)add;
                                    rows_any &= rows_any - 1;
                                }
                            }
                        }
                    }
                    if (rem <= 128u) break;
                    pos += 128; pc += 128; rem -= 128;
#pragma unroll
                    for (int k = 0; k < 4; k++) x[k] = xn[k];
                }
                return false;
            };
            // row r = `row` (fixed per call) of one finished tile: matrix cell / candidate edge for column j = cb*32+lane
            auto emit_row = [&](uint32_t row, uint32_t cb, uint32_t c) {
                const uint32_t i = rb * 32 + row, j = cb * 32 + lane;
                if (a.mode == KH_MODE_MATRIX) {
                    if (i >= a.i0 && i < a.i1 && j >= a.j0 && j < a.j1)
                        a.out[(size_t)(i - a.i0) * (a.j1 - a.j0) + (j - a.j0)] = i == j ? s_len[row] : c;   // private ranks are not in the table
                    return;
                }
                bool keep = false;
                if (c >= 1u && i < a.n && i >= a.row_min && j < i) {
                    const uint32_t s0 = s_len[row], s1 = a.size[j];
                    const uint32_t mn = min(s0, s1), mx = max(s0, s1);
                    keep = mn > 0 && !((long long)mx > (long long)a.radio * (long long)mn);   // MST.cpp:1481-1487
                    if (a.uf && keep) { kh_uf_edge(a, i, j, c, s0, s1); keep = false; }
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                if (ballot) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(&a.counters[0], (unsigned long long)__popc(ballot));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (keep) {
                        const unsigned long long idx = base + __popc(ballot & ((1u << lane) - 1));
                        if (idx < a.edge_cap) {
                            tc_pair e; e.i = i; e.j = j; e.common = c;
                            a.edges[idx] = e;
                        } else {
                            a.counters[1] = 1ULL;
                        }
                    }
                }
            };
            // one tile (row block x column block cb) on this warp
            auto process_tile = [&](const uint32_t cb) {
                const uint32_t D = a.bu_count[cb * KH_DREC];
                const unsigned long long lbase = a.off[cb * 32];
                const uint32_t *keys = a.bu_keys + lbase, *cmasks = a.bu_mask + lbase;
                uint32_t lo = 0, hi = D;
                if (P > 1) {
                    lo = p == 0 ? 0u : kh_lower_bound(keys, 0, D, piv_lo);
                    hi = (p + 1 == P) ? D : kh_lower_bound(keys, lo, D, piv_hi);
                }
                uint32_t *tst = state + (size_t)(cb - c0 / 32) * KJ_TILE_STATE;
                uint32_t anyhit = p == 0 ? 0u : tst[KJ_TILE_WORDS];
                if (anyhit) {
#pragma unroll
                    for (int q = 0; q < 16; q++) cntw[lane + 32 * q] = tst[lane + 32 * q];
                } else {
#pragma unroll
                    for (int q = 0; q < 16; q++) cntw[lane + 32 * q] = 0u;
                }
                __syncwarp();
                const bool deferred = stream_slice(keys, cmasks, lo, hi, unit_defers, anyhit);
                __syncwarp();
                if (deferred) {                                   // redone by the whole CTA below
                    uint32_t slot = 0;
                    if (lane == 0) slot = atomicAdd(&s_nheavy, 1u);
                    slot = __shfl_sync(0xffffffffu, slot, 0);
                    if (slot < KJ_HEAVY_CAP) {
                        if (lane == 0) s_heavy[slot] = cb;
                        return;
                    }
                    // list full: this warp redoes the tile on its own
                    anyhit = p == 0 ? 0u : tst[KJ_TILE_WORDS];
                    if (anyhit) {
#pragma unroll
                        for (int q = 0; q < 16; q++) cntw[lane + 32 * q] = tst[lane + 32 * q];
                    } else {
#pragma unroll
                        for (int q = 0; q < 16; q++) cntw[lane + 32 * q] = 0u;
                    }
                    __syncwarp();
                    stream_slice(keys, cmasks, lo, hi, false, anyhit);
                    __syncwarp();
                }
                if (p + 1 < P) {
                    if (lane == 0) tst[KJ_TILE_WORDS] = anyhit;
                    if (anyhit) {
#pragma unroll
                        for (int q = 0; q < 16; q++) tst[lane + 32 * q] = cntw[lane + 32 * q];
                    }
                } else if (a.mode == KH_MODE_MATRIX || anyhit) {
                    for (uint32_t r = 0; r < 32; r++) emit_row(r, cb, anyhit ? (uint32_t)cnt[r * 32 + lane] : 0u);
                }
                __syncwarp();
            };
            // hand-out: 8 column blocks per grab, only the flagged ones are visited
            for (;;) {
                uint32_t g = 0;
                if (lane == 0) g = atomicAdd(&s_next, 1u);
                g = __shfl_sync(0xffffffffu, g, 0);
                const uint32_t cb0 = cb_begin + g * 8u;
                if (cb0 >= cb_end) break;
                uint32_t bits = 0xFFu;
                if (flag_row) bits = __funnelshift_r(__ldg(flag_row + (cb0 >> 5)), __ldg(flag_row + (cb0 >> 5) + 1), cb0 & 31) & 0xFFu;
                if (cb_end - cb0 < 8u) bits &= (1u << (cb_end - cb0)) - 1u;
                while (bits) {
                    const uint32_t cb = cb0 + (uint32_t)__ffs(bits) - 1u;
                    bits &= bits - 1u;
                    process_tile(cb);
                }
            }
            __syncthreads();
            // ---- deferred tiles: all warps share one tile -- warp w streams the w-th 32nd of the list into its own
            //      counters, then sums row w over the 32 counter sets and finishes that row
            const uint32_t nheavy = min(s_nheavy, KJ_HEAVY_CAP);
            for (uint32_t hx = 0; hx < nheavy; hx++) {
                const uint32_t cb = s_heavy[hx];
                const uint32_t D = a.bu_count[cb * KH_DREC];
                const unsigned long long lbase = a.off[cb * 32];
                const uint32_t *keys = a.bu_keys + lbase, *cmasks = a.bu_mask + lbase;
                uint32_t lo = 0, hi = D;
                if (P > 1) {
                    lo = p == 0 ? 0u : kh_lower_bound(keys, 0, D, piv_lo);
                    hi = (p + 1 == P) ? D : kh_lower_bound(keys, lo, D, piv_hi);
                }
                uint32_t *tst = state + (size_t)(cb - c0 / 32) * KJ_TILE_STATE;
                const uint32_t prev = p == 0 ? 0u : tst[KJ_TILE_WORDS];
                if (warp == 0 && prev) {
#pragma unroll
                    for (int q = 0; q < 16; q++) cntw[lane + 32 * q] = tst[lane + 32 * q];
                } else {
#pragma unroll
                    for (int q = 0; q < 16; q++) cntw[lane + 32 * q] = 0u;
                }
                __syncwarp();
                const uint32_t chunk = ((hi - lo + 31) / 32 + 31) / 32 * 32;      // whole groups of 32 keys per warp
                const uint32_t slo = min(hi, lo + (uint32_t)warp * chunk), shi = min(hi, slo + chunk);
                uint32_t anyhit = 0;
                stream_slice(keys, cmasks, slo, shi, false, anyhit);
                __syncthreads();
                uint32_t total = 0;
#pragma unroll 8
                for (int v = 0; v < 32; v++) total += reinterpret_cast<const uint16_t *>(cnt_all + v * KJ_TILE_WORDS)[warp * 32 + lane];
                if (p + 1 < P) {
                    const uint32_t other = __shfl_down_sync(0xffffffffu, total, 1);
                    if (!(lane & 1)) tst[warp * 16 + (lane >> 1)] = (total & 0xFFFFu) | (other << 16);
                    if (tid == 0) tst[KJ_TILE_WORDS] = 1u;
                } else {
                    emit_row((uint32_t)warp, cb, total);
                }
                __syncthreads();
            }
            __syncthreads();
            if (tid < 32) s_lo[tid] = s_hi[tid];
        }
        __syncthreads();
    }
}

// =====================================================================================================
// K3S  k3s_tile_kernel -- the flagged tiles of sparsely related row blocks, one tile per CTA, union lists staged with TMA.
//
// With the tile flags (k3t_tileflags.cu) a row block of a family-structured set shares ranks with a handful of column
// blocks.  K3J still pays a table build, a scan of the flag row and CTA-wide barriers per row block and runs the few
// dense tiles of a unit one after the other (ncu: `barrier` the top stall, 16 us per tile).  Here the work item is the
// flagged TILE: a CTA of 8 warps stages the two union lists of the tile -- ascending distinct ranks + 32-bit member
// masks of the row block and of the column block, ~1,300 entries each for related sketches -- into shared memory with
// four TMA bulk copies (cp.async.bulk.shared::cluster.global, completion on an mbarrier), intersects them (the column
// keys are looked up in the row list by a branch-free binary search in shared memory; both lists are sorted and
// duplicate-free), and accumulates count[r][c] += popc(rows holding key & columns holding key) exactly as K3J does.
// No hash table, no partitions (a list that does not fit is the host's cue to route the row to K3J), three CTAs per SM
// overlap each other's copies.  Same epilogue (candidate edge / union-find) and the same `common` for every pair.
constexpr uint32_t KS_THREADS = 256, KS_WARPS = KS_THREADS / 32;
constexpr uint32_t KS_RCAP = 4096;                                   // row-block list entries per stage
constexpr uint32_t KS_CCHUNK = 2048;                                 // column list entries per stage
constexpr uint32_t KS_ROW_MAX_TILES = 0xFFFFFFFFu;                            // rows with more flagged tiles stay on K3J (table amortised); TCLUST_KS_ROW_MAX
constexpr size_t KS_SMEM = (size_t)(KS_RCAP + 4) * 8 + (size_t)(KS_CCHUNK + 4) * 8 + (size_t)KS_WARPS * KJ_TILE_WORDS * 4 + 64;

__device__ __forceinline__ void ks_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ks_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ks_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// TMA bulk copy global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on the mbarrier
__device__ __forceinline__ void ks_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// 32 x 32 bit transpose across a warp: in, lane k holds word k; out, lane j holds bit j of every word (bit k = word k's
// bit j).  Five butterfly stages of one shuffle each (instead of 32 ballots).
__device__ __forceinline__ uint32_t ks_transpose32(uint32_t x, uint32_t lane)
{
#pragma unroll
    for (int st = 0; st < 5; st++) {
        const uint32_t s = 16u >> st;
        const uint32_t mask = st == 0 ? 0x0000FFFFu : st == 1 ? 0x00FF00FFu : st == 2 ? 0x0F0F0F0Fu : st == 3 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        x = (lane & s) ? ((x & ~mask) | ((y >> s) & mask)) : ((x & mask) | ((y << s) & ~mask));
    }
    return x;
}

// row `row` of a finished tile (rb, cb): candidate edge / union for column j = cb*32 + lane (the K3J epilogue with the
// row sizes read from global memory)
__device__ __forceinline__ void ks_emit_row(const KHArgs &a, uint32_t rb, uint32_t row, uint32_t cb, uint32_t lane, uint32_t c)
{
    const uint32_t i = rb * 32 + row, j = cb * 32 + lane;
    bool keep = false;
    if (c >= 1u && i < a.n && i >= a.row_min && j < i) {
        const uint32_t s0 = a.size[i], s1 = a.size[j];
        const uint32_t mn = min(s0, s1), mx = max(s0, s1);
        keep = mn > 0 && !((long long)mx > (long long)a.radio * (long long)mn);   // MST.cpp:1481-1487
        if (a.uf && keep) { kh_uf_edge(a, i, j, c, s0, s1); keep = false; }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (ballot) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&a.counters[0], (unsigned long long)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            const unsigned long long idx = base + __popc(ballot & ((1u << lane) - 1));
            if (idx < a.edge_cap) {
                tc_pair e; e.i = i; e.j = j; e.common = c;
                a.edges[idx] = e;
            } else {
                a.counters[1] = 1ULL;
            }
        }
    }
}

__global__ void __launch_bounds__(KS_THREADS, 3) k3s_tile_kernel(const KHArgs a, const uint2 *tiles, const uint32_t *ntiles_dev,
                                                               uint32_t t_begin, uint32_t t_end)
{
    extern __shared__ __align__(128) uint8_t ks_smem[];
    uint32_t *rkeys = reinterpret_cast<uint32_t *>(ks_smem);
    uint32_t *rmask = rkeys + (KS_RCAP + 4);
    uint32_t *ckeys = rmask + (KS_RCAP + 4);
    uint32_t *cmask = ckeys + (KS_CCHUNK + 4);
    uint32_t *cnt_all = cmask + (KS_CCHUNK + 4);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(cnt_all + KS_WARPS * KJ_TILE_WORDS);
    __shared__ uint32_t s_tile;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_r = (uint32_t)__cvta_generic_to_shared(&bars[0]), bar_c = (uint32_t)__cvta_generic_to_shared(&bars[1]);
    const uint32_t rkeys_s = (uint32_t)__cvta_generic_to_shared(rkeys), rmask_s = (uint32_t)__cvta_generic_to_shared(rmask);
    const uint32_t ckeys_s = (uint32_t)__cvta_generic_to_shared(ckeys), cmask_s = (uint32_t)__cvta_generic_to_shared(cmask);
    uint32_t *cntw = cnt_all + warp * KJ_TILE_WORDS;
    uint16_t *cnt = reinterpret_cast<uint16_t *>(cntw);
    if (tid == 0) {
        ks_mbar_init(bar_r, 1); ks_mbar_init(bar_c, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t par_r = 0, par_c = 0;
    const uint32_t t_lim = min(t_end, *ntiles_dev);

    for (;;) {
        if (tid == 0) s_tile = t_begin + (uint32_t)atomicAdd(&a.counters[2], 1ULL);
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= t_lim) break;
        const uint2 tl = tiles[t];
        const uint32_t rb = tl.x, cb = tl.y;
        const uint32_t Dr = a.bu_count[rb * KH_DREC], Dc = a.bu_count[cb * KH_DREC];
        const bool live = Dr > 0 && Dc > 0 && rb * 32 + 31 >= a.row_min;
        const unsigned long long rel = a.off[rb * 32], cel = a.off[cb * 32];
        const uint32_t lead_r = (uint32_t)(rel & 3ull), lead_c = (uint32_t)(cel & 3ull);   // bulk copies start 16-byte aligned
        // count[r][lane] of this warp's share of the keys lives in registers (no shared-memory read-modify-write chain)
        uint32_t acc[32];
#pragma unroll
        for (int r = 0; r < 32; r++) acc[r] = 0u;
        uint32_t anyhit = 0;
        const uint32_t *rk = rkeys + lead_r, *rm = rmask + lead_r, *ck = ckeys + lead_c, *cm = cmask + lead_c;
        // Both lists are ascending: they are streamed through shared memory in stages (KS_RCAP row entries, KS_CCHUNK
        // column entries) like the two inputs of a merge -- the stage whose last key is smaller is replaced by its
        // successor -- so every entry is copied once whatever the list lengths.
        uint32_t r0 = 0, c0 = 0;
        bool need_r = live, need_c = live;
        while (live && r0 < Dr && c0 < Dc) {
            const uint32_t rlen = min(KS_RCAP, Dr - r0), len = min(KS_CCHUNK, Dc - c0);
            if (tid == 0) {
                if (need_r) {
                    const uint32_t bytes = ((lead_r + rlen + 3u) & ~3u) * 4u;
                    ks_mbar_expect_tx(bar_r, 2u * bytes);
                    ks_bulk_g2s(rkeys_s, a.bu_keys + (rel - lead_r + r0), bytes, bar_r);
                    ks_bulk_g2s(rmask_s, a.bu_mask + (rel - lead_r + r0), bytes, bar_r);
                }
                if (need_c) {
                    const uint32_t bytes = ((lead_c + len + 3u) & ~3u) * 4u;
                    ks_mbar_expect_tx(bar_c, 2u * bytes);
                    ks_bulk_g2s(ckeys_s, a.bu_keys + (cel - lead_c + c0), bytes, bar_c);
                    ks_bulk_g2s(cmask_s, a.bu_mask + (cel - lead_c + c0), bytes, bar_c);
                }
            }
            if (need_r) { ks_mbar_wait(bar_r, par_r); par_r ^= 1u; }
            if (need_c) { ks_mbar_wait(bar_c, par_c); par_c ^= 1u; }
            const uint32_t rlo = rk[0], rhi = rk[rlen - 1], clo = ck[0], chi = ck[len - 1];
            if (!(chi < rlo || clo > rhi)) {
                uint32_t top = 1;                                    // largest power of two <= rlen
                while (top * 2 <= rlen) top *= 2;
                for (uint32_t g = warp * 32; g < len; g += KS_WARPS * 32) {
                    const uint32_t k = g + lane;
                    const uint32_t x = k < len ? ck[k] : 0xFFFFFFFFu;
                    uint32_t mk = 0u;
                    if (__any_sync(0xffffffffu, x >= rlo && x <= rhi)) {
                        // number of row keys < x: branch-free lower bound, the same trip count for every lane
                        uint32_t lo = 0;
                        for (uint32_t step = top; step; step >>= 1) {
                            const uint32_t nx = lo + step;
                            if (nx <= rlen && rk[nx - 1] < x) lo = nx;
                        }
                        if (lo < rlen && rk[lo] == x) mk = rm[lo];
                    }
                    if (!__any_sync(0xffffffffu, mk != 0u)) continue;
                    anyhit = 1u;
                    const uint32_t cmk = mk ? cm[k] : 0u;
                    // lane k holds (rows holding key k, columns holding key k); transposed: lane r holds the keys row r
                    // holds, lane c the keys column c holds; count[r][c] += popc(keys of row r & keys of column c)
                    const uint32_t rowkeys = ks_transpose32(mk, lane), colkeys = ks_transpose32(cmk, lane);
#pragma unroll
                    for (int r = 0; r < 32; r++) acc[r] += __popc(__shfl_sync(0xffffffffu, rowkeys, r) & colkeys);
                }
            }
            __syncthreads();                  // every warp is done with both stages
            need_r = rhi <= chi; need_c = chi <= rhi;
            if (need_r) r0 += rlen;
            if (need_c) c0 += len;
        }
        const bool tile_hit = __syncthreads_or((int)anyhit) != 0;
        if (tile_hit) {
#pragma unroll
            for (int r = 0; r < 32; r++) cnt[r * 32 + lane] = (uint16_t)acc[r];
        }
        __syncthreads();
        if (tile_hit) {
            // warp w finishes rows 4w .. 4w+3: sum over the 8 warps' counter sets, then the epilogue
#pragma unroll 1
            for (uint32_t r = warp * 4; r < warp * 4 + 4; r++) {
                uint32_t total = 0;
#pragma unroll
                for (uint32_t v = 0; v < KS_WARPS; v++) total += reinterpret_cast<const uint16_t *>(cnt_all + v * KJ_TILE_WORDS)[r * 32 + lane];
                ks_emit_row(a, rb, r, cb, lane, total);
            }
        }
        __syncthreads();
    }
}

// flagged tiles of the units routed to K3S, compacted into a device list (one warp per unit; slots reserved with one
// atomic per 1024 column blocks)
__global__ void k3s_list_kernel(const tc_kh_unit *units, uint32_t nunits, const uint32_t *flags, uint32_t words, uint2 *tiles,
                                uint32_t *ntiles, uint32_t cap)
{
    const uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (u >= nunits) return;
    const tc_kh_unit un = units[u];
    const uint32_t cb0 = un.c0 / 32, cb1 = (un.c1 + 31) / 32;       // column blocks [cb0, cb1)
    const uint32_t *row = flags + (size_t)un.rb * words;
    for (uint32_t w0 = cb0 >> 5; w0 * 32 < cb1; w0 += 32) {
        const uint32_t wi = w0 + lane;
        uint32_t v = wi * 32 < cb1 ? row[wi] : 0u;
        if (wi == (cb0 >> 5) && (cb0 & 31)) v &= ~((1u << (cb0 & 31)) - 1u);
        if (wi * 32 + 32 > cb1 && wi * 32 < cb1) v &= (cb1 & 31) ? ((1u << (cb1 & 31)) - 1u) : 0xFFFFFFFFu;
        const uint32_t pc = __popc(v);
        uint32_t incl = pc;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += y; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (!total) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(ntiles, total);
        base = __shfl_sync(0xffffffffu, base, 0) + incl - pc;
        while (v) {
            const uint32_t b = (uint32_t)__ffs(v) - 1u;
            if (base < cap) tiles[base] = make_uint2(un.rb, wi * 32 + b);
            base++;
            v &= v - 1u;
        }
    }
}

// ---- per-upload setup: partition pivots of every row block
int tc_kh_prepare(tc_ctx *ctx)
{
    const uint32_t nrb = (ctx->n + 31) / 32;
    cudaStream_t st = ctx->stream;
    // ---- distinct counts and pivot candidates of every row block ([rb * KH_DREC] = distinct count of block rb,
    //      0xFFFFFFFF: no union list); the buffers only grow (cudaMalloc/cudaFree stall for milliseconds now and then)
    if ((uint64_t)nrb * KH_DREC > ctx->cap_bu_count) {
        cudaFree(ctx->d_bu_count); ctx->d_bu_count = nullptr; ctx->cap_bu_count = 0;
        TC_CUDA(cudaMalloc(&ctx->d_bu_count, (size_t)nrb * KH_DREC * 4));
        ctx->cap_bu_count = (uint64_t)nrb * KH_DREC;
    }
    uint32_t *d_dist = ctx->d_bu_count;
    TC_CUDA(cudaMemsetAsync(d_dist, 0, (size_t)nrb * KH_DREC * 4, st));
    if (!(ctx->attr_set & 1u)) {        // per context: the attribute is a per-device property
        TC_CUDA(cudaFuncSetAttribute(k3h_distinct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KH_SORT_CAP * 5)));
        ctx->attr_set |= 1u;
    }
    // union lists of the 32-sketch blocks (K3J): as many slots as elements
    if (ctx->total + 8 > ctx->cap_bu) {      // + 8: K3S copies whole 16-byte quads, up to 3 entries past a list's end
        cudaFree(ctx->d_bu_keys); cudaFree(ctx->d_bu_mask);
        ctx->d_bu_keys = nullptr; ctx->d_bu_mask = nullptr; ctx->cap_bu = 0;
        if (cudaMalloc(&ctx->d_bu_keys, (ctx->total + 8) * 4) != cudaSuccess || cudaMalloc(&ctx->d_bu_mask, (ctx->total + 8) * 4) != cudaSuccess) {
            tc_set_error("tc_kh_prepare: out of device memory for the block union lists"); return 1;
        }
        ctx->cap_bu = ctx->total + 8;
    }
    TC_CHECK(ctx->d_multi, "tc_kh_prepare: no multiplicity bitmap");
    const char *sort_env = getenv("TCLUST_DISTINCT_SORT");       // "plain": one pass per stride (the first version)
    const int blocked = !(sort_env && !strcmp(sort_env, "plain"));
    const char *dist_env = getenv("TCLUST_DISTINCT");            // "sort": the block sort for every block (round 1)
    const bool hash_first = !(dist_env && !strcmp(dist_env, "sort"));
    std::vector<uint32_t> dist((size_t)nrb * KH_DREC);
    cudaError_t e = cudaSuccess;
    if (hash_first) {
        if (!(ctx->attr_set & 8u)) {
            TC_CUDA(cudaFuncSetAttribute(k3h_distinct_hash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KDH_SMEM));
            ctx->attr_set |= 8u;
        }
        k3h_distinct_hash_kernel<<<nrb, KH_THREADS, KDH_SMEM, st>>>(ctx->d_ranks, reinterpret_cast<unsigned long long *>(ctx->d_off),
                                                                   ctx->d_size, ctx->n, ctx->d_multi, d_dist, ctx->d_bu_keys, ctx->d_bu_mask);
        ctx->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(dist.data(), d_dist, dist.size() * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { tc_set_error("tc_kh_prepare: %s", cudaGetErrorString(e)); return 1; }
    }
    // blocks the hash kernel could not hold (more than KDH_CAP distinct shared ranks), or all of them: the block sort
    std::vector<uint32_t> redo;
    for (uint32_t rb = 0; rb < nrb; rb++)
        if (!hash_first || dist[(size_t)rb * KH_DREC] == KDH_RETRY) redo.push_back(rb);
    if (!redo.empty()) {
        uint32_t *d_list = nullptr;
        if (redo.size() < nrb) {
            void *ws = nullptr;
            if (tc_scratch(ctx, redo.size() * 4, &ws)) return 1;
            d_list = static_cast<uint32_t *>(ws);
            TC_CUDA(cudaMemcpyAsync(d_list, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, st));
        }
        k3h_distinct_kernel<<<(unsigned)redo.size(), KH_THREADS, KH_SORT_CAP * 5, st>>>(
            ctx->d_ranks, reinterpret_cast<unsigned long long *>(ctx->d_off), ctx->d_size, ctx->n, ctx->d_multi, d_dist, ctx->d_bu_keys,
            ctx->d_bu_mask, blocked, d_list);
        ctx->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(dist.data(), d_dist, dist.size() * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { tc_set_error("tc_kh_prepare: %s", cudaGetErrorString(e)); return 1; }
    }
    // ---- partitions: by distinct keys where known, else by the multiset total
    std::vector<uint32_t> pivoff(nrb + 1, 0);
    bool any_fallback = false;
    for (uint32_t rb = 0; rb < nrb; rb++) {
        uint64_t keys = dist[(size_t)rb * KH_DREC];
        if (keys == 0xFFFFFFFFu) {
            any_fallback = true;
            keys = 0;
            for (uint32_t l = 0; l < 32 && rb * 32 + l < ctx->n; l++) keys += ctx->h_size[rb * 32 + l];
        }
        const uint32_t P = (uint32_t)std::max<uint64_t>(1, (keys + (KH_KCAP - 64) - 1) / (KH_KCAP - 64));
        pivoff[rb + 1] = pivoff[rb] + (P - 1);
    }
    std::vector<uint32_t> piv(std::max<size_t>(pivoff[nrb], 1), 0u);
    for (uint32_t rb = 0; rb < nrb; rb++) {
        const uint32_t *d = &dist[(size_t)rb * KH_DREC];
        const uint32_t P = pivoff[rb + 1] - pivoff[rb] + 1;
        if (d[0] == 0xFFFFFFFFu || P == 1) continue;
        // candidates: (2,1) (3,1) (3,2) (4,1) (4,2) (4,3)
        const uint32_t first = P == 2 ? 1 : P == 3 ? 2 : 4;
        for (uint32_t q = 0; q + 1 < P; q++) piv[pivoff[rb] + q] = d[first + q];
    }
    ctx->bu_count.assign(nrb, 0xFFFFFFFFu);
    ctx->kh_parts.assign(nrb, 1);
    for (uint32_t rb = 0; rb < nrb; rb++) {
        ctx->bu_count[rb] = dist[(size_t)rb * KH_DREC];
        ctx->kh_parts[rb] = pivoff[rb + 1] - pivoff[rb] + 1;
    }
    int rc = 1;
    do {
        if (nrb + 1 > ctx->cap_pivoff) {
            cudaFree(ctx->d_kh_pivoff); ctx->d_kh_pivoff = nullptr; ctx->cap_pivoff = 0;
            if (cudaMalloc(&ctx->d_kh_pivoff, (size_t)(nrb + 1) * 4) != cudaSuccess) { tc_set_error("tc_kh_prepare: out of device memory"); break; }
            ctx->cap_pivoff = nrb + 1;
        }
        if (piv.size() > ctx->cap_pivots) {
            cudaFree(ctx->d_kh_pivots); ctx->d_kh_pivots = nullptr; ctx->cap_pivots = 0;
            if (cudaMalloc(&ctx->d_kh_pivots, piv.size() * 4) != cudaSuccess) { tc_set_error("tc_kh_prepare: out of device memory"); break; }
            ctx->cap_pivots = piv.size();
        }
        if (cudaMemcpyAsync(ctx->d_kh_pivoff, pivoff.data(), (size_t)(nrb + 1) * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(ctx->d_kh_pivots, piv.data(), piv.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            tc_set_error("tc_kh_prepare: %s", cudaGetErrorString(cudaGetLastError())); break;
        }
        if (any_fallback && pivoff[nrb] > 0) {
            const unsigned blocks = (unsigned)(((uint64_t)nrb * 32 + 255) / 256);
            k3h_pivot_kernel<<<blocks, 256, 0, st>>>(ctx->d_ranks, reinterpret_cast<unsigned long long *>(ctx->d_off), ctx->d_size,
                                                     ctx->n, ctx->d_kh_pivoff, nrb, d_dist, ctx->d_kh_pivots);
            ctx->launches++;
        }
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {   // host vectors outlive the copies
            tc_set_error("tc_kh_prepare: %s", cudaGetErrorString(cudaGetLastError())); break;
        }
        rc = 0;
    } while (0);
    if (!rc) rc = tc_tile_flags_build(ctx, dist);
    return rc;
}

// edge_list: `units` is ctx->kh_edge_units, which stays resident on the device between passes (no per-pass copy
// from pageable host memory in front of the kernel)
static int kh_launch(tc_ctx *ctx, KHArgs &a, const std::vector<tc_kh_unit> &units_all, bool edge_list = false,
                     bool block_kernel = false, bool record_start = true, const tc_kh_unit *d_resident = nullptr,
                     size_t range_begin = 0, size_t range_end = ~(size_t)0)
{
    // units [range_begin, range_end) of a device-resident list (bounded-memory passes launch a list piece by piece)
    struct Range { size_t b, e; size_t size() const { return e - b; } bool empty() const { return e <= b; } };
    const Range units{d_resident ? std::min(range_begin, units_all.size()) : 0, d_resident ? std::min(range_end, units_all.size()) : units_all.size()};
    if (d_resident) d_resident += units.b;
    if (!(ctx->attr_set & 2u)) {
        TC_CUDA(cudaFuncSetAttribute(k3h_join_kernel<KH_MODE_EDGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KH_SMEM));
        TC_CUDA(cudaFuncSetAttribute(k3h_join_kernel<KH_MODE_CAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KH_SMEM));
        TC_CUDA(cudaFuncSetAttribute(k3j_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KJ_SMEM));
        ctx->attr_set |= 2u;
    }
    if (units.empty()) return 0;
    if (!d_resident && units.size() > ctx->kh_units_cap) {
        cudaFree(ctx->d_kh_units); ctx->d_kh_units = nullptr; ctx->kh_units_cap = 0; ctx->kh_units_on_device = false;
        TC_CUDA(cudaMalloc(&ctx->d_kh_units, units.size() * sizeof(tc_kh_unit)));
        ctx->kh_units_cap = units.size();
    }
    if (!d_resident && !(edge_list && ctx->kh_units_on_device)) {
        TC_CUDA(cudaMemcpyAsync(ctx->d_kh_units, units_all.data(), units.size() * sizeof(tc_kh_unit), cudaMemcpyHostToDevice,
                                ctx->stream));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));   // `units` (host) must outlive the copy
        ctx->kh_units_on_device = edge_list;
    }
    const bool nosync = d_resident && ctx->pass_nosync;       // nothing on the host has to outlive a copy
    const long long mine = ((long long)units.size() - a.part + a.nparts - 1) / a.nparts;
    if (mine <= 0) { if (!nosync) TC_CUDA(cudaStreamSynchronize(ctx->stream)); return 0; }
    const int grid = (int)std::min<long long>(mine, ctx->num_sms);
    if (!ctx->d_kh_state) TC_CUDA(cudaMalloc(&ctx->d_kh_state, (size_t)ctx->num_sms * KH_STATE_WORDS_MAX * 4));
    a.ranks = ctx->d_ranks; a.off = reinterpret_cast<const unsigned long long *>(ctx->d_off); a.size = ctx->d_size;
    a.n = ctx->n; a.pivots = ctx->d_kh_pivots; a.pivoff = ctx->d_kh_pivoff;
    a.units = d_resident ? d_resident : ctx->d_kh_units; a.nunits = (uint32_t)units.size();
    a.counters = ctx->d_edge_count; a.state = ctx->d_kh_state;
    a.bu_keys = ctx->d_bu_keys; a.bu_mask = ctx->d_bu_mask; a.bu_count = ctx->d_bu_count; a.multi = ctx->d_multi;
    a.tile_flags = nullptr; a.tile_words = 0;
    if (a.mode == KH_MODE_EDGES && ctx->tile_words) { a.tile_flags = ctx->d_tile_flags; a.tile_words = ctx->tile_words; }
    a.uf = nullptr; a.uf_perm = nullptr; a.uf_jmin = 2.0;
    if (a.mode == KH_MODE_EDGES && ctx->uf_parent) { a.uf = ctx->uf_parent; a.uf_perm = ctx->uf_perm; a.uf_jmin = ctx->uf_jmin; a.containment = ctx->uf_containment; }
    TC_CUDA(cudaMemsetAsync(ctx->d_edge_count + 2, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->ev_pass0 && record_start) cudaEventRecord(ctx->ev_pass0, ctx->stream);   // kernel-only duration, tc_last_pass_kernel_ms
    if (block_kernel) k3j_block_kernel<<<grid, KH_THREADS, KJ_SMEM, ctx->stream>>>(a);
    else if (a.mode == KH_MODE_CAND) k3h_join_kernel<KH_MODE_CAND><<<grid, KH_THREADS, KH_SMEM, ctx->stream>>>(a);
    else k3h_join_kernel<KH_MODE_EDGES><<<grid, KH_THREADS, KH_SMEM, ctx->stream>>>(a);     // matrix and edge epilogues
    if (ctx->ev_pass1) cudaEventRecord(ctx->ev_pass1, ctx->stream);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    if (!nosync) TC_CUDA(cudaStreamSynchronize(ctx->stream));   // `units` (host) must outlive the copy
    return 0;
}

int tc_launch_join_matrix(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *d_out)
{
    std::vector<tc_kh_unit> units;
    for (uint32_t c0 = j0; c0 < j1; c0 += KH_CSEG)
        for (uint32_t rb = i0 / 32; rb * 32 < i1; rb++) units.push_back(tc_kh_unit{rb, c0, std::min(j1, c0 + KH_CSEG)});
    KHArgs a{};
    a.mode = KH_MODE_MATRIX; a.part = 0; a.nparts = 1;
    a.out = d_out; a.i0 = i0; a.i1 = i1; a.j0 = j0; a.j1 = j1;
    return kh_launch(ctx, a, units);
}

// unit list of an edge pass: pairs (i >= start_index, j < i); row block rb needs the columns [0, rb*32+31); units are
// segment-major so that all CTAs stream the same column segment at the same time (L2 reuse)
static void kh_edge_unit_list(tc_ctx *ctx, uint32_t start_index, uint32_t cseg)
{
    const uint32_t n = ctx->n, nrb = (n + 31) / 32;
    if (ctx->kh_edge_units_n == n && ctx->kh_edge_units_start == start_index && ctx->kh_edge_units_cseg == cseg) return;
    ctx->kh_edge_units.clear();
    for (uint32_t c0 = 0; c0 < n; c0 += cseg)
        for (uint32_t rb = std::max(c0 / 32, start_index / 32); rb < nrb; rb++) {
            const uint32_t c1 = std::min<uint32_t>(std::min(n, c0 + cseg), rb * 32 + 31);
            if (c1 > c0) ctx->kh_edge_units.push_back(tc_kh_unit{rb, c0, c1});
        }
    ctx->kh_edge_units_n = n; ctx->kh_edge_units_start = start_index; ctx->kh_edge_units_cseg = cseg;
    ctx->kh_units_on_device = false; ctx->kj_split_valid = false;
}

int tc_launch_join_edges(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, uint32_t start_index)
{
    kh_edge_unit_list(ctx, start_index, KH_CSEG);
    KHArgs a{};
    a.mode = KH_MODE_EDGES; a.part = part; a.nparts = nparts ? nparts : 1; a.row_min = start_index;
    a.edges = ctx->d_edges; a.edge_cap = ctx->edge_cap; a.radio = radio;
    return kh_launch(ctx, a, ctx->kh_edge_units, true);
}

// greedy candidates of the query batch [q0, q1) (q0 multiple of 32) on the hash-join kernel: against the representative
// table (vs_reps; columns = ctx->d_rep_ids) or against the batch itself (earlier members only).  Appends to
// ctx->d_cands; tc_cand.first_hit carries the smallest shared RANK (the reference orders `touched` by the first
// shared query hash, greedy.cpp:1148-1164; ranks are order-preserving, so comparing them is the same order).
int tc_launch_join_cands(tc_ctx *ctx, uint32_t q0, uint32_t q1, int vs_reps, int cmin_fixed, int containment, double jmin)
{
    const uint32_t ccount = vs_reps ? ctx->nrep : q1, cfirst = vs_reps ? 0u : q0;
    std::vector<tc_kh_unit> units;
    for (uint32_t c0 = cfirst; c0 < ccount; c0 += KH_CSEG)
        for (uint32_t rb = q0 / 32; rb * 32 < q1; rb++) {
            uint32_t c1 = std::min(ccount, c0 + KH_CSEG);
            if (!vs_reps) c1 = std::min<uint32_t>(c1, rb * 32 + 31);      // only columns with an id below some row of the block
            if (c1 > c0) units.push_back(tc_kh_unit{rb, c0, c1});
        }
    if (units.empty()) return 0;
    KHArgs a{};
    a.mode = KH_MODE_CAND; a.part = 0; a.nparts = 1;
    a.col_ids = vs_reps ? ctx->d_rep_ids : nullptr; a.q0 = q0; a.q1 = q1;
    a.cands = ctx->d_cands; a.edge_cap = ctx->cand_cap; a.lower_only = vs_reps ? 0 : 1;
    a.cmin_fixed = cmin_fixed; a.containment = containment; a.jmin = jmin; a.one_plus_jmin = 1.0 + jmin;
    return kh_launch(ctx, a, units);
}

// ---- K3J launchers.  A unit goes to the block kernel when every column block it touches has a union list; the rest
// (blocks whose 32 sketches hold more than KH_SORT_CAP elements) stays on K3H.
static bool kj_unit_listed(const tc_ctx *ctx, const tc_kh_unit &u)
{
    for (uint32_t cb = u.c0 / 32; cb * 32 < u.c1; cb++)
        if (cb >= ctx->bu_count.size() || ctx->bu_count[cb] == 0xFFFFFFFFu) return false;
    return true;
}

int tc_launch_blockjoin_matrix(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *d_out)
{
    std::vector<tc_kh_unit> blk, old;
    for (uint32_t c0 = j0 / 32 * 32; c0 < j1; c0 += KH_CSEG)
        for (uint32_t rb = i0 / 32; rb * 32 < i1; rb++) {
            const tc_kh_unit u{rb, c0, std::min(j1, c0 + KH_CSEG)};
            (kj_unit_listed(ctx, u) ? blk : old).push_back(u);
        }
    KHArgs a{};
    a.mode = KH_MODE_MATRIX; a.part = 0; a.nparts = 1;
    a.out = d_out; a.i0 = i0; a.i1 = i1; a.j0 = j0; a.j1 = j1;
    if (kh_launch(ctx, a, blk, false, true)) return 1;
    for (tc_kh_unit &u : old) u.c0 = std::max(u.c0, j0);
    return kh_launch(ctx, a, old);
}

// K3J pays when the sketches are related to their neighbours: the row block fits one table partition and the union
// lists of the unit's column blocks are clearly shorter than the sketches themselves (measured at N = 100k, s = 1000:
// related neighbours 86 ms vs 153 ms on K3H; unrelated neighbours 657 ms vs 292 ms).  Prefix sums over the column
// blocks make the test O(1) per unit (40k units x 128 blocks at N = 100k).
struct KjRouting {
    std::vector<uint64_t> keys;        // prefix sum of the union-list lengths
    std::vector<uint32_t> unlisted;    // prefix count of blocks without a list
    explicit KjRouting(const tc_ctx *ctx)
    {
        const uint32_t nb = (ctx->n + 31) / 32;
        keys.assign(nb + 1, 0); unlisted.assign(nb + 1, 0);
        for (uint32_t b = 0; b < nb; b++) {
            const bool listed = b < ctx->bu_count.size() && ctx->bu_count[b] != 0xFFFFFFFFu;
            keys[b + 1] = keys[b] + (listed ? ctx->bu_count[b] : 0u);
            unlisted[b + 1] = unlisted[b] + (listed ? 0u : 1u);
        }
    }
    bool listed(const tc_kh_unit &u) const
    {
        const uint32_t b0 = u.c0 / 32, b1 = (u.c1 + 31) / 32;
        return b1 < unlisted.size() && unlisted[b1] == unlisted[b0];
    }
    bool pays(const tc_ctx *ctx, const tc_kh_unit &u) const
    {
        // a row block that shares ranks with few column blocks visits few tiles on the block kernel (tile flags)
        if (ctx->tile_words && u.rb < ctx->tile_rowcount.size() && (uint64_t)ctx->tile_rowcount[u.rb] * 8 <= u.rb + 1) return true;
        if (ctx->kh_parts.empty() || ctx->kh_parts[u.rb] > 2) return false;
        const uint32_t b0 = u.c0 / 32, b1 = (u.c1 + 31) / 32;
        const uint64_t elems = ctx->h_off[std::min<uint32_t>(ctx->n, b1 * 32)] - ctx->h_off[b0 * 32];
        return (keys[b1] - keys[b0]) * 10 <= elems * 6;
    }
};

// columns per K3J unit: the widest power of two up to KJ_CSEG_WIDEST that still leaves every CTA of every rank 2 units.
// Since the union lists hold shared ranks only they fit the L2 as a whole (32 MB at N = 100k), so nothing is gained from
// all CTAs streaming the same column segment, while every unit pays a table build and a barrier tail: 8.92 ms per pass
// at 16384 columns, 7.68 at 32768, 7.21 at 65536 (tools/unit_width_sweep.py).  At N = 100k a unit is a whole row of
// tiles; tc_launch_blockjoin_edges hands the widest units out first.
// (mirrored by rabbittclust_b200/mst_dist.py: join_cseg)
static uint32_t kj_pick_cseg(uint32_t n, uint32_t start_index, uint32_t nparts, int num_sms)
{
    const uint64_t nrb = (n + 31) / 32, rb0 = start_index / 32;
    uint64_t per_cta = 2;
    if (const char *e = getenv("TCLUST_KJ_MIN_UNITS")) per_cta = std::max(1, atoi(e));   // tuning aid (tools/unit_width_sweep.py)
    uint32_t cseg_max = KJ_CSEG_WIDEST;
    if (const char *e = getenv("TCLUST_KJ_CSEG_MAX")) cseg_max = (uint32_t)std::min<long>(KJ_CSEG_WIDEST, std::max<long>(KH_CSEG, atol(e)));
    for (uint32_t cseg = cseg_max; cseg > KH_CSEG; cseg >>= 1) {
        uint64_t units = 0;
        for (uint64_t c0 = 0; c0 < n; c0 += cseg) units += nrb - std::max<uint64_t>(c0 / 32, rb0);
        if (units >= per_cta * nparts * (uint64_t)num_sms) return cseg;
    }
    return KH_CSEG;
}

// tiles [t0, t1) of this context's sparse tile list on K3S
static int ks_launch(tc_ctx *ctx, int radio, uint32_t start_index, uint32_t t0, uint32_t t1, bool record_start)
{
    if (ctx->ks_ntiles_host == 0 || t0 >= ctx->ks_ntiles_host) return 0;
    if (!(ctx->attr_set & 4u)) {
        TC_CUDA(cudaFuncSetAttribute(k3s_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_SMEM));
        ctx->attr_set |= 4u;
    }
    KHArgs a{};
    a.mode = KH_MODE_EDGES; a.part = 0; a.nparts = 1; a.row_min = start_index;
    a.edges = ctx->d_edges; a.edge_cap = ctx->edge_cap; a.radio = radio;
    a.ranks = ctx->d_ranks; a.off = reinterpret_cast<const unsigned long long *>(ctx->d_off); a.size = ctx->d_size; a.n = ctx->n;
    a.counters = ctx->d_edge_count;
    a.bu_keys = ctx->d_bu_keys; a.bu_mask = ctx->d_bu_mask; a.bu_count = ctx->d_bu_count; a.multi = ctx->d_multi;
    a.uf = nullptr; a.uf_perm = nullptr; a.uf_jmin = 2.0;
    if (ctx->uf_parent) { a.uf = ctx->uf_parent; a.uf_perm = ctx->uf_perm; a.uf_jmin = ctx->uf_jmin; a.containment = ctx->uf_containment; }
    const uint32_t cnt = std::min(t1, ctx->ks_ntiles_host) - t0;
    const int grid = (int)std::min<uint64_t>(cnt, (uint64_t)ctx->num_sms * 3);
    TC_CUDA(cudaMemsetAsync(ctx->d_edge_count + 2, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->ev_pass0 && record_start) cudaEventRecord(ctx->ev_pass0, ctx->stream);
    k3s_tile_kernel<<<grid, KS_THREADS, KS_SMEM, ctx->stream>>>(a, static_cast<const uint2 *>(ctx->d_ks_tiles), ctx->d_ks_count, t0, t1);
    if (ctx->ev_pass1) cudaEventRecord(ctx->ev_pass1, ctx->stream);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    if (!ctx->pass_nosync) TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int tc_launch_blockjoin_edges(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, uint32_t start_index, bool force_block, bool plan_only)
{
    // the unit list (also what tc_allpairs_pass counts pairs from), split by kernel
    kh_edge_unit_list(ctx, start_index, kj_pick_cseg(ctx->n, start_index, nparts ? nparts : 1, ctx->num_sms));
    // the routing of this rank's units is decided once per (sketch set, start_index, part) and kept on the device
    const uint64_t key = ((uint64_t)part << 40) ^ ((uint64_t)(nparts ? nparts : 1) << 33) ^ ((uint64_t)start_index << 1) ^ (force_block ? 1u : 0u);
    if (!ctx->kj_split_valid || ctx->kj_split_key != key) {
        std::vector<tc_kh_unit> &blk = ctx->kj_units_blk, &old = ctx->kj_units_old, &sps = ctx->ks_units;
        blk.clear(); old.clear(); sps.clear();
        const KjRouting route(ctx);
        // K3S takes the units of row blocks that share ranks with few column blocks (tile flags) and whose own list can be
        // staged; sets with a block too large for a union list keep everything on K3J / K3H (TCLUST_K3S=0: never)
        const char *ks_env = getenv("TCLUST_K3S");
        const bool ks_on = ctx->tile_words && !force_block && route.unlisted.back() == 0 && !(ks_env && !strcmp(ks_env, "0"));
        uint64_t ks_cap = 0;
        uint32_t ks_row_max = KS_ROW_MAX_TILES;          // (measured: K3S wins in the all-dense regime too, 3.1 vs 8.8 ms on a one-species set)
        if (const char *e = getenv("TCLUST_KS_ROW_MAX")) ks_row_max = (uint32_t)std::max(0l, atol(e));
        for (size_t u = part; u < ctx->kh_edge_units.size(); u += (nparts ? nparts : 1)) {
            const tc_kh_unit &un = ctx->kh_edge_units[u];
            if (ks_on && un.rb < ctx->tile_rowcount.size() && ctx->tile_rowcount[un.rb] < ks_row_max && ctx->bu_count[un.rb] != 0xFFFFFFFFu) {
                sps.push_back(un);
                ks_cap += std::min<uint64_t>(ctx->tile_rowcount[un.rb], (un.c1 - un.c0 + 31) / 32 + 1);
                continue;
            }
            if (route.listed(un) && (force_block || route.pays(ctx, un))) {
                // a partitioned row block keeps 2 KB of state per tile between its partitions: units of KJ_CSEG_MAX columns
                const uint32_t w = ctx->kh_parts[un.rb] > 1 ? KJ_CSEG_MAX : un.c1 - un.c0;
                for (uint32_t c0 = un.c0; c0 < un.c1; c0 += w) blk.push_back(tc_kh_unit{un.rb, c0, std::min(un.c1, c0 + w)});
                continue;
            }
            for (uint32_t c0 = un.c0; c0 < un.c1; c0 += KH_CSEG)          // K3H's per-CTA column state holds KH_CSEG columns
                old.push_back(tc_kh_unit{un.rb, c0, std::min(un.c1, c0 + KH_CSEG)});
        }
        // widest units first: the dynamic hand-out then ends on the small ones (a unit is up to a whole row of tiles)
        // (with tile flags the work of a unit is its flagged tiles: the share of the row's flagged tiles its width covers)
        auto weight = [&](const tc_kh_unit &x) -> uint64_t {
            const uint64_t w = x.c1 - x.c0;
            if (!ctx->tile_words || x.rb >= ctx->tile_rowcount.size()) return w;
            return (uint64_t)ctx->tile_rowcount[x.rb] * 32 * w / ((uint64_t)x.rb * 32 + 32) + 1;
        };
        std::stable_sort(blk.begin(), blk.end(), [&](const tc_kh_unit &x, const tc_kh_unit &y) { return weight(x) > weight(y); });
        const size_t all_units = blk.size() + old.size() + sps.size();
        if (all_units > ctx->kj_units_cap || !ctx->d_kj_units) {
            cudaFree(ctx->d_kj_units);
            ctx->d_kj_units = nullptr; ctx->kj_units_cap = 0;
            TC_CUDA(cudaMalloc(&ctx->d_kj_units, std::max<size_t>(all_units, 1) * sizeof(tc_kh_unit)));
            ctx->kj_units_cap = all_units;
        }
        if (!blk.empty()) TC_CUDA(cudaMemcpy(ctx->d_kj_units, blk.data(), blk.size() * sizeof(tc_kh_unit), cudaMemcpyHostToDevice));
        if (!old.empty()) TC_CUDA(cudaMemcpy(ctx->d_kj_units + blk.size(), old.data(), old.size() * sizeof(tc_kh_unit), cudaMemcpyHostToDevice));
        ctx->ks_ntiles_host = 0;
        if (!sps.empty()) {
            // the tile list of the sparse rows: filled on the device from the flag rows, its length stays on the device
            tc_kh_unit *d_sps = ctx->d_kj_units + blk.size() + old.size();
            TC_CUDA(cudaMemcpy(d_sps, sps.data(), sps.size() * sizeof(tc_kh_unit), cudaMemcpyHostToDevice));
            if (ks_cap + 1 > ctx->cap_ks_tiles || !ctx->d_ks_tiles) {
                cudaFree(ctx->d_ks_tiles); ctx->d_ks_tiles = nullptr; ctx->cap_ks_tiles = 0;
                TC_CUDA(cudaMalloc(&ctx->d_ks_tiles, (ks_cap + 1) * sizeof(uint2)));
                ctx->cap_ks_tiles = ks_cap + 1;
            }
            if (!ctx->d_ks_count) TC_CUDA(cudaMalloc(&ctx->d_ks_count, 2 * sizeof(uint32_t)));
            TC_CUDA(cudaMemsetAsync(ctx->d_ks_count, 0, 2 * sizeof(uint32_t), ctx->stream));
            k3s_list_kernel<<<(unsigned)((sps.size() * 32 + 255) / 256), 256, 0, ctx->stream>>>(
                d_sps, (uint32_t)sps.size(), ctx->d_tile_flags, ctx->tile_words, static_cast<uint2 *>(ctx->d_ks_tiles), ctx->d_ks_count,
                (uint32_t)std::min<uint64_t>(ks_cap + 1, 0xFFFFFFFFull));
            ctx->launches++;
            TC_CUDA(cudaGetLastError());
            uint32_t h_cnt = 0;
            TC_CUDA(cudaMemcpyAsync(&h_cnt, ctx->d_ks_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
            TC_CUDA(cudaStreamSynchronize(ctx->stream));
            TC_CHECK(h_cnt <= ks_cap + 1, "tc_allpairs_pass: sparse tile list overflow (%u > %llu)", h_cnt, (unsigned long long)ks_cap + 1);
            ctx->ks_ntiles_host = h_cnt;
        }
        ctx->kj_split_key = key; ctx->kj_split_valid = true;
    }
    if (plan_only) return 0;
    if (ks_launch(ctx, radio, start_index, 0, 0xFFFFFFFFu, true)) return 1;
    KHArgs a{};
    a.mode = KH_MODE_EDGES; a.part = 0; a.nparts = 1; a.row_min = start_index;   // the lists already hold this part only
    a.edges = ctx->d_edges; a.edge_cap = ctx->edge_cap; a.radio = radio;
    const bool first = ctx->ks_ntiles_host == 0;
    if (kh_launch(ctx, a, ctx->kj_units_blk, false, true, first, ctx->d_kj_units)) return 1;
    return kh_launch(ctx, a, ctx->kj_units_old, false, false, first && ctx->kj_units_blk.empty(), ctx->d_kj_units + ctx->kj_units_blk.size());
}

// ---- the same pass piece by piece (bounded edge memory, tclust_api.cu mst_bounded), after
// tc_launch_blockjoin_edges(..., plan_only = true): items [u0, u1) of list 0 = sparse tiles (K3S), 1 = block-join units
// (K3J), 2 = hash-join units (K3H)
size_t tc_blockjoin_list_size(const tc_ctx *ctx, int list)
{
    return list == 0 ? ctx->ks_ntiles_host : list == 1 ? ctx->kj_units_blk.size() : ctx->kj_units_old.size();
}

int tc_blockjoin_run_range(tc_ctx *ctx, int radio, uint32_t start_index, int list, size_t u0, size_t u1)
{
    if (list == 0) return ks_launch(ctx, radio, start_index, (uint32_t)u0, (uint32_t)u1, true);
    KHArgs a{};
    a.mode = KH_MODE_EDGES; a.part = 0; a.nparts = 1; a.row_min = start_index;
    a.edges = ctx->d_edges; a.edge_cap = ctx->edge_cap; a.radio = radio;
    if (list == 1) return kh_launch(ctx, a, ctx->kj_units_blk, false, true, true, ctx->d_kj_units, u0, u1);
    return kh_launch(ctx, a, ctx->kj_units_old, false, false, true, ctx->d_kj_units + ctx->kj_units_blk.size(), u0, u1);
}

// most candidate edges item u of that list can emit: a tile 1024; a unit 32 rows x its columns, or 1024 per flagged tile
// of a whole-row unit
uint64_t tc_blockjoin_unit_bound(const tc_ctx *ctx, int list, size_t u)
{
    if (list == 0) return 1024;
    const tc_kh_unit &x = list == 1 ? ctx->kj_units_blk[u] : ctx->kj_units_old[u];
    uint64_t b = (uint64_t)32 * (x.c1 - x.c0);
    if (ctx->tile_words && x.c0 == 0 && x.c1 + 32 > x.rb * 32 && x.rb < ctx->tile_rowcount.size())
        b = std::min<uint64_t>(b, (uint64_t)ctx->tile_rowcount[x.rb] * 1024);
    return b;
}
