// k3_tile.cu -- K3: all-pairs sorted-sketch intersection, the HP-2 hot kernel.
//
// Replaces the candidate loop of compute_minhash_mst (reference src/MST.cpp:1409-1435,
// 1467-1533) and the pairwise merge of RabbitSketch's u64_intersect_vector_* kernels
// (RabbitSketch/src/MinHash.cpp:1199-1643) with a dense, data-independent pass:
// every pair (i, j) gets its exact |A_i n A_j| (semantic A).
//
// Design (see DESIGN.md "K3"):
//  * sketches are stored as order-preserving dense u32 ranks, 32 sketches interleaved per
//    group: gt[group][row][lane] -- element `row` of sketch `lane` of the group.
//  * one CTA (1024 threads) owns a 32x32 tile: 32 row sketches and 32 column sketches live in
//    shared memory in the same interleaved layout, so sketch x occupies shared-memory bank x.
//    Thread (warp w, lane l) merges row sketch l with column sketch (l+w)&31: in every warp
//    the 32 row reads hit 32 distinct banks and so do the 32 column reads -> every LDS is
//    conflict-free although each lane walks its own two pointers.
//  * a tile is processed in value-range chunks: each chunk loads up to CAP elements of every
//    stream, cuts all 64 streams at a common pivot value (the smallest "last loaded" element
//    of the streams that did not fit), merges, and continues behind the pivot.  Sketches of
//    any length are handled; 1000-element sketches take two chunks.
//  * the merge step is branch-free: both pointers advance by predicated adds; the match count
//    is recovered from the pointer advance (advances - steps).  Sentinels terminate a window.
#include "tc_common.cuh"

namespace {

constexpr int K3_THREADS = 1024;
constexpr int K3_CAP = 888;          // window rows per stream and chunk
constexpr int K3_U = 8;              // merge steps between termination checks == sentinel rows
constexpr int K3_ROWS = K3_CAP + K3_U;

struct K3Meta {
    unsigned long long gbase[64];  // element offset of stream x (element 0) in gt
    uint32_t len[64];
    uint32_t pos[64];
    uint32_t cut[64];
    uint32_t pivot;
    uint32_t more;                 // bit0: some row stream has data left, bit1: some column stream
    long long tile;
};
constexpr size_t K3_SMEM = (size_t)2 * K3_ROWS * 32 * 4 + sizeof(K3Meta);

enum { K3_MODE_MATRIX = 0, K3_MODE_EDGES = 1, K3_MODE_CAND = 2 };

// a set of sketches in the group-interleaved layout: member m = group*32 + lane
struct K3Set {
    const uint32_t *gt;              // interleaved ranks
    const unsigned long long *goff;  // ngroups+1 element offsets into gt
    const uint32_t *len;             // length of member m
    const uint32_t *ids;             // global sketch id of member m (nullptr: identity)
    uint32_t n, ngroups;
};

struct K3Args {
    K3Set row, col;
    // tile space: row groups [rg0, rg1), col groups [cg0, cg1); triangular => only cg <= rg
    uint32_t rg0, rg1, cg0, cg1;
    int triangular;
    long long ntiles;
    uint32_t part, nparts;
    unsigned long long *counters;  // [0] edge/candidate count, [1] overflow, [2] tile counter
    // matrix mode (validation): common (and first_hit when FH) for rows [i0,i1) x col members [j0,j1)
    uint32_t *out; uint32_t *out_fh; uint32_t i0, i1, j0, j1;
    // edge mode (MST candidate graph)
    tc_pair *edges; unsigned long long edge_cap; int radio;
    // candidate mode (greedy representative scan): emit (query, rep, common, first_hit) when
    // common >= 1 and common + 1 >= common_min(pair)  (host applies the exact rule)
    tc_cand *cands; int lower_only;   // lower_only: emit only col id < row id (intra-batch tiles)
    int cmin_fixed;                   // >= 0: fixed-size Mash mode; < 0: per-pair rule below
    int containment; double jmin, one_plus_jmin;
    // semantic B (SEMB kernels): the merge of a pair stops after semb_size union elements
    // (MinHash::jaccard, RabbitSketch/src/MinHash.cpp:982-1101; scalar form :1199-1229)
    uint32_t semb_size;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void tile_to_rc(const K3Args &a, long long t, uint32_t &rg, uint32_t &cg)
{
    if (a.triangular) {
        // t = r(r+1)/2 + c, 0 <= c <= r   (r, c relative to rg0 == cg0)
        long long r = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while (r * (r + 1) / 2 > t) r--;
        while ((r + 1) * (r + 2) / 2 <= t) r++;
        rg = a.rg0 + (uint32_t)r;
        cg = a.cg0 + (uint32_t)(t - r * (r + 1) / 2);
    } else {
        uint32_t ncg = a.cg1 - a.cg0;
        rg = a.rg0 + (uint32_t)(t / ncg);
        cg = a.cg0 + (uint32_t)(t % ncg);
    }
}

template <int MODE, bool FH, bool SEMB>
__global__ void __launch_bounds__(K3_THREADS, 1) k3_tile_kernel(const K3Args a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t *sA = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *sB = sA + K3_ROWS * 32;
    K3Meta *m = reinterpret_cast<K3Meta *>(sB + K3_ROWS * 32);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int prow = lane, pcol = (lane + warp) & 31;
    const uint32_t sA_addr = (uint32_t)__cvta_generic_to_shared(sA);
    const uint32_t sB_addr = (uint32_t)__cvta_generic_to_shared(sB);

    // sentinel rows behind the windows never change
    if (tid < K3_U * 32) {
        sA[K3_CAP * 32 + tid] = TC_SENT_A;
        sB[K3_CAP * 32 + tid] = TC_SENT_B;
    }

    for (;;) {
        if (tid == 0) {
            unsigned long long k = atomicAdd(&a.counters[2], 1ULL);
            m->tile = (long long)(k * a.nparts + a.part);
        }
        __syncthreads();
        const long long t = m->tile;
        if (t >= a.ntiles) break;
        uint32_t rg, cg;
        tile_to_rc(a, t, rg, cg);
        if (tid < 64) {
            const K3Set &S = tid < 32 ? a.row : a.col;
            const uint32_t g = tid < 32 ? rg : cg;
            const uint32_t sk = g * TC_GROUP + (tid & 31);
            m->len[tid] = (g < S.ngroups && sk < S.n) ? S.len[sk] : 0u;
            m->pos[tid] = 0;
            m->gbase[tid] = (g < S.ngroups ? S.goff[g] : 0ULL) + (tid & 31);
        }
        uint32_t acc = 0;
        uint32_t first_hit = 0xFFFFFFFFu;   // FH: index in the row sketch of the first shared element
        uint32_t budget = SEMB ? a.semb_size : 0u;   // SEMB: union elements this pair may still consume
        __syncthreads();

        for (;;) {  // ---- value-range chunks ----
            {   // load: warp w fills rows w, w+32, ... of both regions; lane = stream
                const uint32_t pA = m->pos[lane], lA = m->len[lane];
                const uint32_t pB = m->pos[32 + lane], lB = m->len[32 + lane];
                const uint32_t *gA = a.row.gt + m->gbase[lane];
                const uint32_t *gB = a.col.gt + m->gbase[32 + lane];
#pragma unroll 4
                for (int r = warp; r < K3_CAP; r += 32) {
                    const uint32_t eA = pA + r, eB = pB + r;
                    uint32_t vA = TC_SENT_A, vB = TC_SENT_B;
                    if (eA < lA) vA = __ldg(gA + (size_t)eA * 32);
                    if (eB < lB) vB = __ldg(gB + (size_t)eB * 32);
                    sA[r * 32 + lane] = vA;
                    sB[r * 32 + lane] = vB;
                }
                if (tid == 0) { m->pivot = TC_RANK_MAX; m->more = 0; }
            }
            __syncthreads();
            if (tid < 64) {  // pivot = smallest last-loaded element among streams that did not fit
                const uint32_t *col = (tid < 32 ? sA : sB) + (tid & 31);
                if (m->pos[tid] + K3_CAP < m->len[tid]) atomicMin(&m->pivot, col[(K3_CAP - 1) * 32]);
            }
            __syncthreads();
            if (tid < 64) {  // cut every stream at the pivot, drop a sentinel run behind the cut
                uint32_t *col = (tid < 32 ? sA : sB) + (tid & 31);
                const uint32_t piv = m->pivot;
                const uint32_t rem = m->len[tid] - m->pos[tid];
                uint32_t lo = 0, hi = rem < (uint32_t)K3_CAP ? rem : (uint32_t)K3_CAP;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (col[mid * 32] <= piv) lo = mid + 1; else hi = mid;
                }
                m->cut[tid] = lo;
                const uint32_t s = tid < 32 ? TC_SENT_A : TC_SENT_B;
                const uint32_t end = lo + K3_U < (uint32_t)K3_CAP ? lo + K3_U : (uint32_t)K3_CAP;
                for (uint32_t r = lo; r < end; r++) col[r * 32] = s;
                if (m->pos[tid] + lo < m->len[tid]) atomicOr(&m->more, tid < 32 ? 1u : 2u);
            }
            __syncthreads();
            if (!SEMB) {   // ---- branch-free two-pointer merge, one pair per thread ----
                const uint32_t pa0 = sA_addr + prow * 4, pb0 = sB_addr + pcol * 4;
                uint32_t pa = pa0, pb = pb0;
                uint32_t va = lds_u32(pa), vb = lds_u32(pb);
                uint32_t nblk = 0;
                uint32_t fh_addr = 0xFFFFFFFFu;
                while (max(va, vb) < TC_SENT_B) {
#pragma unroll
                    for (int u = 0; u < K3_U; u++) {
                        const bool le = va <= vb, ge = va >= vb;
                        if (FH) { if (le && ge) fh_addr = min(fh_addr, pa); }
                        if (le) pa += 128;
                        if (ge) pb += 128;
                        va = lds_u32(pa);
                        vb = lds_u32(pb);
                    }
                    nblk++;
                }
                acc += (((pa - pa0) + (pb - pb0)) >> 7) - nblk * K3_U;
                if (FH) {
                    if (first_hit == 0xFFFFFFFFu && fh_addr != 0xFFFFFFFFu)
                        first_hit = m->pos[prow] + ((fh_addr - pa0) >> 7);
                }
            } else {
                // ---- semantic B: the same merge with a step budget.  Every active step consumes exactly one
                // element of the union in ascending order (steps taken when both streams sit on their
                // sentinels are spurious: they advance pb and the step counter alike and end the chunk).
                const uint32_t pa0 = sA_addr + prow * 4, pb0 = sB_addr + pcol * 4;
                uint32_t pa = pa0, pb = pb0;
                uint32_t va = lds_u32(pa), vb = lds_u32(pb);
                uint32_t steps = 0, left = budget;
                while (max(va, vb) < TC_SENT_B && left > 0) {
#pragma unroll
                    for (int u = 0; u < K3_U; u++) {
                        const bool act = (uint32_t)u < left;
                        const bool le = act && va <= vb, ge = act && va >= vb;
                        if (le) pa += 128;
                        if (ge) pb += 128;
                        va = lds_u32(pa);
                        vb = lds_u32(pb);
                    }
                    const uint32_t d = min(left, (uint32_t)K3_U);
                    steps += d;
                    left -= d;
                }
                const uint32_t cc = (((pa - pa0) + (pb - pb0)) >> 7) - steps;
                acc += cc;
                if (max(va, vb) >= TC_SENT_B) {
                    // one stream of this value range is used up: the rest of the other stream (all <= pivot)
                    // precedes the next chunk in the union, so the whole chunk counts against the budget
                    const uint32_t used = m->cut[prow] + m->cut[32 + pcol] - cc;
                    budget = budget > used ? budget - used : 0u;
                } else {
                    budget = 0;      // stopped inside the chunk: sketch_size union elements consumed
                }
            }
            const uint32_t more = m->more;
            __syncthreads();
            if (more != 3u) break;  // all row streams or all column streams are exhausted
            if (tid < 64) m->pos[tid] += m->cut[tid];
            __syncthreads();
        }

        // ---- epilogue ----
        const uint32_t i = rg * TC_GROUP + prow, jm = cg * TC_GROUP + pcol;   // member indices
        if (MODE == K3_MODE_MATRIX) {
            if (i >= a.i0 && i < a.i1 && jm >= a.j0 && jm < a.j1) {
                const size_t o = (size_t)(i - a.i0) * (a.j1 - a.j0) + (jm - a.j0);
                a.out[o] = acc;
                if (FH) a.out_fh[o] = first_hit;
            }
        } else {
            bool keep = false;
            const uint32_t s0 = m->len[prow], s1 = m->len[32 + pcol];
            uint32_t j = jm;
            if (MODE == K3_MODE_EDGES) {
                if (SEMB) {
                    keep = i < a.row.n && j < i && acc >= 1u;   // modifyMST: no size-ratio filter
                } else if (i < a.row.n && j < i && acc >= 1u) {
                    const uint32_t mn = min(s0, s1), mx = max(s0, s1);
                    // reference: skip if max_size > radio * min_size (int arithmetic), MST.cpp:1481-1487
                    keep = mn > 0 && !((long long)mx > (long long)a.radio * (long long)mn);
                }
            } else {
                if (i < a.row.n && jm < a.col.n && acc >= 1u) {
                    if (a.col.ids) j = a.col.ids[jm];
                    int cmin = a.cmin_fixed;
                    if (SEMB) {
                        // J_B = common / min(S, |A u B|) <= common / min(S, max size): conservative bound,
                        // the host applies MinHash::distance exactly
                        cmin = (int)ceil(__dmul_rn(a.jmin, (double)min(a.semb_size, max(s0, s1))));
                    } else if (cmin < 0) {   // greedy.cpp:1211-1219, IEEE-exact operations in the host's order
                        const double t = a.containment
                            ? __dmul_rn(a.jmin, (double)min(s0, s1))
                            : __ddiv_rn(__dmul_rn(a.jmin, (double)(s0 + s1)), a.one_plus_jmin);
                        cmin = (int)ceil(t);
                    }
                    keep = (long long)acc + 1 >= (long long)cmin && (!a.lower_only || j < i);
                }
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            if (ballot) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&a.counters[0], (unsigned long long)__popc(ballot));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) {
                    const unsigned long long idx = base + __popc(ballot & ((1u << lane) - 1));
                    if (idx < a.edge_cap) {
                        if (MODE == K3_MODE_EDGES) {
                            tc_pair e; e.i = i; e.j = j; e.common = acc;
                            a.edges[idx] = e;
                        } else {
                            tc_cand c; c.query = i; c.rep = j; c.common = acc; c.first_hit = first_hit;
                            a.cands[idx] = c;
                        }
                    } else {
                        a.counters[1] = 1ULL;
                    }
                }
            }
        }
    }
}

// independent validation kernel: one thread per pair, two-pointer merge on the ORIGINAL u64
// hashes straight from global memory (checks the rank transform and the tiled kernel).
__global__ void pair_common_ref_kernel(const uint64_t *hashes, const unsigned long long *off, uint32_t i0,
                                       uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *out)
{
    const uint32_t nj = j1 - j0;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)(i1 - i0) * nj) return;
    const uint32_t i = i0 + (uint32_t)(idx / nj), j = j0 + (uint32_t)(idx % nj);
    const uint64_t *pa = hashes + off[i], *ea = hashes + off[i + 1];
    const uint64_t *pb = hashes + off[j], *eb = hashes + off[j + 1];
    uint32_t c = 0;
    while (pa < ea && pb < eb) {
        const uint64_t x = *pa, y = *pb;
        c += x == y;
        pa += x <= y;
        pb += x >= y;
    }
    out[idx] = c;
}

template <int MODE, bool FH, bool SEMB>
int k3_launch_t(tc_ctx *ctx, const K3Args &a, int grid)
{
    // set on every launch: the attribute is per device (a process may hold contexts on several) and the call is host-only
    TC_CUDA(cudaFuncSetAttribute(k3_tile_kernel<MODE, FH, SEMB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K3_SMEM));
    k3_tile_kernel<MODE, FH, SEMB><<<grid, K3_THREADS, K3_SMEM, ctx->stream>>>(a);
    return 0;
}

K3Set main_set(const tc_ctx *ctx)
{
    K3Set s{};
    s.gt = ctx->d_gt; s.goff = reinterpret_cast<const unsigned long long *>(ctx->d_goff);
    s.len = ctx->d_size; s.ids = nullptr; s.n = ctx->n; s.ngroups = ctx->ngroups;
    return s;
}

K3Set rep_set(const tc_ctx *ctx)
{
    K3Set s{};
    s.gt = ctx->d_rep_gt; s.goff = reinterpret_cast<const unsigned long long *>(ctx->d_rep_goff);
    s.len = ctx->d_rep_len; s.ids = ctx->d_rep_ids; s.n = ctx->nrep; s.ngroups = (ctx->nrep + TC_GROUP - 1) / TC_GROUP;
    return s;
}

int k3_launch(tc_ctx *ctx, K3Args &a, int mode, bool fh)
{
    a.counters = ctx->d_edge_count;
    TC_CUDA(cudaMemsetAsync(ctx->d_edge_count + 2, 0, sizeof(unsigned long long), ctx->stream));
    long long my_tiles = (a.ntiles - a.part + a.nparts - 1) / a.nparts;
    if (my_tiles <= 0) return 0;
    int grid = (int)(my_tiles < ctx->num_sms ? my_tiles : ctx->num_sms);
    int rc;
    const bool semb = a.semb_size > 0;
    if (ctx->ev_pass0) cudaEventRecord(ctx->ev_pass0, ctx->stream);      // kernel-only duration, tc_last_pass_kernel_ms
    if (mode == K3_MODE_MATRIX)
        rc = semb ? k3_launch_t<K3_MODE_MATRIX, false, true>(ctx, a, grid)
           : fh   ? k3_launch_t<K3_MODE_MATRIX, true, false>(ctx, a, grid)
                  : k3_launch_t<K3_MODE_MATRIX, false, false>(ctx, a, grid);
    else if (mode == K3_MODE_EDGES)
        rc = semb ? k3_launch_t<K3_MODE_EDGES, false, true>(ctx, a, grid) : k3_launch_t<K3_MODE_EDGES, false, false>(ctx, a, grid);
    else
        rc = semb ? k3_launch_t<K3_MODE_CAND, false, true>(ctx, a, grid) : k3_launch_t<K3_MODE_CAND, true, false>(ctx, a, grid);
    if (rc) return rc;
    if (ctx->ev_pass1) cudaEventRecord(ctx->ev_pass1, ctx->stream);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int tc_launch_tiles_matrix(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *d_out,
                           uint32_t semb_size)
{
    if (tc_ensure_interleaved(ctx)) return 1;
    K3Args a{};
    a.semb_size = semb_size;
    a.row = main_set(ctx); a.col = main_set(ctx);
    a.rg0 = i0 / TC_GROUP; a.rg1 = (i1 + TC_GROUP - 1) / TC_GROUP;
    a.cg0 = j0 / TC_GROUP; a.cg1 = (j1 + TC_GROUP - 1) / TC_GROUP;
    a.triangular = 0;
    a.ntiles = (long long)(a.rg1 - a.rg0) * (a.cg1 - a.cg0);
    a.part = 0; a.nparts = 1;
    a.out = d_out; a.i0 = i0; a.i1 = i1; a.j0 = j0; a.j1 = j1;
    return k3_launch(ctx, a, K3_MODE_MATRIX, false);
}

// queries [q0,q1) of the main set against the representative table: common and first_hit matrices
int tc_launch_tiles_rep_matrix(tc_ctx *ctx, uint32_t q0, uint32_t q1, uint32_t *d_common, uint32_t *d_first_hit)
{
    if (tc_ensure_interleaved(ctx)) return 1;
    K3Args a{};
    a.row = main_set(ctx); a.col = rep_set(ctx);
    a.rg0 = q0 / TC_GROUP; a.rg1 = (q1 + TC_GROUP - 1) / TC_GROUP;
    a.cg0 = 0; a.cg1 = a.col.ngroups;
    a.triangular = 0;
    a.ntiles = (long long)(a.rg1 - a.rg0) * (a.cg1 - a.cg0);
    a.part = 0; a.nparts = 1;
    a.out = d_common; a.out_fh = d_first_hit; a.i0 = q0; a.i1 = q1; a.j0 = 0; a.j1 = ctx->nrep;
    return k3_launch(ctx, a, K3_MODE_MATRIX, true);
}

int tc_launch_tiles_edges(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, uint32_t semb_size)
{
    if (tc_ensure_interleaved(ctx)) return 1;
    K3Args a{};
    a.semb_size = semb_size;
    a.row = main_set(ctx); a.col = main_set(ctx);
    a.rg0 = 0; a.rg1 = ctx->ngroups; a.cg0 = 0; a.cg1 = ctx->ngroups;
    a.triangular = 1;
    a.ntiles = (long long)ctx->ngroups * (ctx->ngroups + 1) / 2;
    a.part = part; a.nparts = nparts ? nparts : 1;
    a.edges = ctx->d_edges; a.edge_cap = ctx->edge_cap; a.radio = radio;
    return k3_launch(ctx, a, K3_MODE_EDGES, false);
}

// greedy candidates of the query batch [q0,q1) (q0 multiple of 32): vs_reps = against the representative
// table, else the batch against itself (col id < row id).  Appends to ctx->d_cands.
int tc_launch_tiles_cands(tc_ctx *ctx, uint32_t q0, uint32_t q1, int vs_reps, int cmin_fixed, int containment,
                          double jmin, uint32_t part, uint32_t nparts, uint32_t semb_size)
{
    if (tc_ensure_interleaved(ctx)) return 1;
    K3Args a{};
    a.semb_size = semb_size;
    a.row = main_set(ctx);
    a.rg0 = q0 / TC_GROUP; a.rg1 = (q1 + TC_GROUP - 1) / TC_GROUP;
    if (vs_reps) {
        a.col = rep_set(ctx);
        a.cg0 = 0; a.cg1 = a.col.ngroups; a.triangular = 0; a.lower_only = 0;
        a.ntiles = (long long)(a.rg1 - a.rg0) * (a.cg1 - a.cg0);
    } else {
        a.col = main_set(ctx);
        a.cg0 = a.rg0; a.cg1 = a.rg1; a.triangular = 1; a.lower_only = 1;
        const long long g = a.rg1 - a.rg0;
        a.ntiles = g * (g + 1) / 2;
    }
    a.part = part; a.nparts = nparts ? nparts : 1;
    a.cands = ctx->d_cands; a.edge_cap = ctx->cand_cap;
    a.cmin_fixed = cmin_fixed; a.containment = containment; a.jmin = jmin; a.one_plus_jmin = 1.0 + jmin;
    a.counters = ctx->d_edge_count;
    return k3_launch(ctx, a, K3_MODE_CAND, true);
}

int tc_launch_pair_common_ref(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *d_out)
{
    const size_t total = (size_t)(i1 - i0) * (j1 - j0);
    if (!total) return 0;
    const int threads = 128;
    pair_common_ref_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, ctx->stream>>>(
        ctx->d_hashes, reinterpret_cast<const unsigned long long *>(ctx->d_off), i0, i1, j0, j1, d_out);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    return 0;
}
