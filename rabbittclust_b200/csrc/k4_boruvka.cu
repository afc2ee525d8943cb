// k4_boruvka.cu -- K4: the whole Boruvka MST on the device, one persistent kernel for all rounds; on several GPUs the
// per-round min-reduction is fused into the same kernel over NVLink peer memory.
//
// Replaces kruskalAlgorithm over the candidate graph of compute_minhash_mst / compute_kssd_mst (src/MST.cpp:59-75,
// 1715-1723; the per-task Kruskal reductions :1373-1386, :571-572, :787).  Kruskal is serial; Boruvka yields the same
// minimum spanning forest under a strict total order of the edges -- here (key, i, j) with key = ~bits(J), J the
// reference's own Jaccard / containment quotient (edge_key below), so "smaller key" == "smaller distance" and the
// sorted multiset of weights and the partition at any threshold equal the reference's.
//
// One cooperative launch (all CTAs co-resident, own grid barrier):
//   round:  init best[] -> phase 1: best_key[c] = min key over the edges leaving component c
//                       -> phase 2: best_edge[c] = min (i<<32|j) among the edges with that key
//                       -> hook: parent[c] = the component at the other end (of a mutual pair the larger label hooks),
//                                the winning edge is appended to the forest
//                       -> chase parents to the roots, relabel comp[]            ... until a round merges nothing.
// Component labels, parents and the forest never leave the device; the host reads back <= n-1 sorted forest entries.
//
// Several GPUs (each holds the candidate edges of its share of the tile space, comp[] replicated): after phase 1
// (and again after phase 2) every GPU has a partial best[] in a buffer its peers can map; after a system-scope barrier
// (epoch flags written into the peers' memory) each GPU pulls the G-1 peer arrays with plain loads over NVLink and
// keeps the element-wise minimum -- an all-reduce(min) of N u64 without leaving the kernel (the NCCL version of the
// same round: kernel + all-reduce + kernel + all-reduce, two host round trips).  Hook / chase / relabel are replicated:
// every GPU ends every round with the same labels and appends the same forest edges.
#include "tc_common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

namespace {

constexpr int K4_THREADS = 512;
constexpr unsigned long long K4_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long K4_CHUNK = 1024;               // candidate edges per liveness flag (one warp, 32 coalesced steps)
constexpr long long K4_SPIN_LIMIT_NS = 10000000000ll;       // a grid barrier that waits longer gives up (error, no hang)
constexpr long long K4_SYS_SPIN_LIMIT_NS = 30000000000ll;   // same for the multi-GPU barrier (a peer may still be in its dense pass)

// similarity key: ~bits(J); J computed exactly as the reference's host code does (MST.cpp:1489-1492:
// (double)common / (size0+size1-common); :1505-1507: / min(size0,size1)).  mode = TC_DIST_*.
__device__ __forceinline__ unsigned long long k4_edge_key(const tc_pair &e, const uint32_t *size, int mode, uint32_t semb)
{
    const uint32_t s0 = size[e.i], s1 = size[e.j];
    const uint32_t uni = s0 + s1 - e.common;
    const uint32_t denom = (mode == 1 || mode == 3) ? min(s0, s1) : mode == 2 ? min(semb, uni) : uni;
    const double J = denom == 0 ? 0.0 : __ddiv_rn((double)e.common, (double)denom);
    return ~(unsigned long long)__double_as_longlong(J);
}

struct K4Less {
    __host__ __device__ bool operator()(const ulonglong2 &x, const ulonglong2 &y) const { return x.x < y.x || (x.x == y.x && x.y < y.y); }
};

struct K4Args {
    const tc_pair *edges; const unsigned long long *edge_counter; unsigned long long edge_cap;   // m = min(*counter, cap)
    const unsigned long long *fin_key, *fin_id; uint32_t nfin;    // extra edges given as (key, i<<32|j) (forest of an earlier batch)
    const uint32_t *size; uint32_t n; int mode; uint32_t semb;
    int32_t *comp; uint32_t *parent;
    unsigned long long *xbuf[8];      // per GPU: best_key[2][xstride], best_edge[2][xstride], flags[8] (double-buffered by round parity)
    size_t xstride;                   // vertices the exchange buffers were laid out for (>= n; the same on every GPU of a group)
    unsigned long long *red;          // 2 x n: reduced key / edge of this round (multi-GPU only)
    int rank, nranks;
    unsigned long long sys_base;      // epoch of the multi-GPU barrier at launch (flags only ever grow: no reset between launches)
    unsigned char *live;              // per chunk of K4_CHUNK candidate edges: 0 = every edge of it lies inside one component (for good)
    ulonglong2 *out;                  // forest: (key, i<<32|j), n entries (preset to all ones)
    uint32_t *ctl;                    // [0] forest count, [1] grid barrier, [2..3] merged per round parity, [4] abort, [5] rounds, [6] some rank's edges overflowed
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_peer_u64(const unsigned long long *p)      // peer memory: never from a stale line
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long k4_now_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// grid barrier: monotone counter, every CTA adds 1 and waits for epoch * gridDim.x; gives up when ctl[4] (abort) is set
__device__ __forceinline__ void k4_grid_sync(uint32_t *ctl, unsigned int &epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(&ctl[1], 1u);
        const long long t0 = k4_now_ns();
        while (ld_acquire_u32(&ctl[1]) < epoch) {
            if (ld_acquire_u32(&ctl[4])) break;
            if (k4_now_ns() - t0 > K4_SPIN_LIMIT_NS) { atomicExch(&ctl[4], 1u); break; }
        }
        __threadfence();
    }
    __syncthreads();
}

// all GPUs: everything written before is visible to the peers after.  Epoch flags: GPU r writes `e` into slot r of every
// peer's flag array and waits until its own array holds e in every slot.
__device__ __forceinline__ void k4_sys_sync(const K4Args &a, unsigned int &epoch, unsigned long long &sys_epoch)
{
    __threadfence_system();
    k4_grid_sync(a.ctl, epoch);
    sys_epoch++;
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)a.nranks) {
        const size_t fl = (size_t)4 * a.xstride;
        st_release_sys_u64(a.xbuf[threadIdx.x] + fl + a.rank, sys_epoch);
        const long long t0 = k4_now_ns();
        while (ld_acquire_sys_u64(a.xbuf[a.rank] + fl + threadIdx.x) < sys_epoch) {
            if (ld_acquire_u32(&a.ctl[4])) break;
            if (k4_now_ns() - t0 > K4_SYS_SPIN_LIMIT_NS) { atomicExch(&a.ctl[4], 2u); break; }
        }
    }
    k4_grid_sync(a.ctl, epoch);
}

__global__ void __launch_bounds__(K4_THREADS) k4_boruvka_kernel(const K4Args a)
{
    const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long gsize = (unsigned long long)gridDim.x * blockDim.x;
    unsigned int epoch = 0;
    unsigned long long sys_epoch = a.sys_base;
    const uint32_t n = a.n;
    unsigned long long m = *a.edge_counter;
    if (m > a.edge_cap) m = a.edge_cap;
    unsigned long long *mine = a.xbuf[a.rank];
    const unsigned long long gwarp = gtid >> 5, nwarps = gsize >> 5, nchunks = (m + K4_CHUNK - 1) / K4_CHUNK;
    const unsigned int lane = threadIdx.x & 31;

    for (unsigned long long v = gtid; v < n; v += gsize) { a.comp[v] = (int32_t)v; a.parent[v] = (uint32_t)v; }
    // a rank whose candidate edges did not fit tells every peer (note slot 8 + rank = this launch's epoch), so that all
    // ranks of the group repeat the call together instead of one of them failing alone
    const bool my_overflow = *a.edge_counter > a.edge_cap || a.edge_counter[1] != 0ull;
    if (a.nranks > 1 && my_overflow && blockIdx.x == 0 && threadIdx.x < (unsigned)a.nranks)
        st_release_sys_u64(a.xbuf[threadIdx.x] + (size_t)4 * a.xstride + 8 + a.rank, a.sys_base + 1);
    if (gtid == 0 && my_overflow) a.ctl[6] = 1u;
    for (uint32_t round = 0;; round++) {
        const uint32_t par = round & 1u;
        unsigned long long *bkey = mine + (size_t)par * a.xstride, *bedge = mine + (size_t)(2 + par) * a.xstride;
        // ---- init
        for (unsigned long long c = gtid; c < n; c += gsize) { bkey[c] = K4_NONE; bedge[c] = K4_NONE; }
        if (gtid == 0) { a.ctl[2 + par] = 0u; a.ctl[5] = round + 1; }
        k4_grid_sync(a.ctl, epoch);
        // ---- phase 1: smallest key leaving every component.  A warp takes a chunk of K4_CHUNK consecutive edges; the pass
        //      appends a tile's edges together, so most chunks lie inside one family and are dead after the first merges:
        //      a dead chunk costs one flag byte instead of 12 KB in every later phase
        for (unsigned long long c = gwarp; c < nchunks; c += nwarps) {
            if (round > 0 && !a.live[c]) continue;
            bool any = false;
            const unsigned long long e1 = min(m, (c + 1) * K4_CHUNK);
            for (unsigned long long e = c * K4_CHUNK + lane; e < e1; e += 64) {       // two edges per lane in flight
                const bool two = e + 32 < e1;
                const tc_pair p = a.edges[e], q = a.edges[two ? e + 32 : e];
                const int32_t ci = a.comp[p.i], cj = a.comp[p.j], di = a.comp[q.i], dj = a.comp[q.j];
                if (ci != cj) {
                    any = true;
                    const unsigned long long k = k4_edge_key(p, a.size, a.mode, a.semb);
                    if (k < bkey[ci]) atomicMin(&bkey[ci], k);
                    if (k < bkey[cj]) atomicMin(&bkey[cj], k);
                }
                if (two && di != dj) {
                    any = true;
                    const unsigned long long k = k4_edge_key(q, a.size, a.mode, a.semb);
                    if (k < bkey[di]) atomicMin(&bkey[di], k);
                    if (k < bkey[dj]) atomicMin(&bkey[dj], k);
                }
            }
            const bool alive = __any_sync(0xffffffffu, any);
            if (lane == 0) a.live[c] = alive ? 1 : 0;
        }
        for (unsigned long long e = gtid; e < a.nfin; e += gsize) {
            const unsigned long long id = a.fin_id[e], k = a.fin_key[e];
            const int32_t ci = a.comp[id >> 32], cj = a.comp[id & 0xFFFFFFFFull];
            if (ci == cj) continue;
            if (k < bkey[ci]) atomicMin(&bkey[ci], k);
            if (k < bkey[cj]) atomicMin(&bkey[cj], k);
        }
        const unsigned long long *rkey = bkey;
        if (a.nranks > 1) {          // all-reduce(min) of best_key over peer memory
            k4_sys_sync(a, epoch, sys_epoch);
            if (round == 0 && gtid < (unsigned long long)a.nranks &&
                ld_acquire_sys_u64(mine + (size_t)4 * a.xstride + 8 + gtid) == a.sys_base + 1) a.ctl[6] = 1u;
            for (unsigned long long c = gtid; c < n; c += gsize) {
                unsigned long long k = bkey[c];
                for (int g = 0; g < a.nranks; g++)
                    if (g != a.rank) k = min(k, ld_peer_u64(a.xbuf[g] + (size_t)par * a.xstride + c));
                a.red[c] = k;
            }
            rkey = a.red;
        }
        k4_grid_sync(a.ctl, epoch);
        // ---- phase 2: smallest (i, j) among the edges that carry a component's best key
        for (unsigned long long c = gwarp; c < nchunks; c += nwarps) {
            if (!a.live[c]) continue;
            const unsigned long long e1 = min(m, (c + 1) * K4_CHUNK);
            for (unsigned long long e = c * K4_CHUNK + lane; e < e1; e += 32) {
                const tc_pair p = a.edges[e];
                const int32_t ci = a.comp[p.i], cj = a.comp[p.j];
                if (ci == cj) continue;
                const unsigned long long k = k4_edge_key(p, a.size, a.mode, a.semb);
                const unsigned long long id = ((unsigned long long)p.i << 32) | p.j;
                if (k == rkey[ci] && id < bedge[ci]) atomicMin(&bedge[ci], id);
                if (k == rkey[cj] && id < bedge[cj]) atomicMin(&bedge[cj], id);
            }
        }
        for (unsigned long long e = gtid; e < a.nfin; e += gsize) {
            const unsigned long long id = a.fin_id[e], k = a.fin_key[e];
            const int32_t ci = a.comp[id >> 32], cj = a.comp[id & 0xFFFFFFFFull];
            if (ci == cj) continue;
            if (k == rkey[ci] && id < bedge[ci]) atomicMin(&bedge[ci], id);
            if (k == rkey[cj] && id < bedge[cj]) atomicMin(&bedge[cj], id);
        }
        const unsigned long long *redge = bedge;
        if (a.nranks > 1) {
            k4_sys_sync(a, epoch, sys_epoch);
            for (unsigned long long c = gtid; c < n; c += gsize) {
                unsigned long long k = bedge[c];
                for (int g = 0; g < a.nranks; g++)
                    if (g != a.rank) k = min(k, ld_peer_u64(a.xbuf[g] + (size_t)(2 + par) * a.xstride + c));
                a.red[(size_t)n + c] = k;
            }
            redge = a.red + n;
        }
        k4_grid_sync(a.ctl, epoch);
        // ---- hook: component c joins the component at the other end of its best edge; when both ends chose the same
        //      edge the larger label hooks (the strict order (key, i, j) admits no longer cycle)
        for (unsigned long long c = gtid; c < n; c += gsize) {
            const unsigned long long id = redge[c];
            if (id == K4_NONE) continue;
            const int32_t ci = a.comp[id >> 32], cj = a.comp[id & 0xFFFFFFFFull];
            const uint32_t other = (uint32_t)((uint32_t)ci == (uint32_t)c ? cj : ci);
            if (redge[other] == id && (uint32_t)c < other) continue;
            a.parent[c] = other;
            const uint32_t slot = atomicAdd(&a.ctl[0], 1u);
            if (slot < n) a.out[slot] = make_ulonglong2(rkey[c], id);
            atomicAdd(&a.ctl[2 + par], 1u);
        }
        k4_grid_sync(a.ctl, epoch);
        if (ld_acquire_u32(&a.ctl[4])) return;               // a barrier gave up
        if (ld_acquire_u32(&a.ctl[2 + par]) == 0u) return;   // nothing merged: the forest is complete
        // ---- roots of the hooked labels, then every vertex takes its root's label
        for (unsigned long long c = gtid; c < n; c += gsize) {
            uint32_t r = a.parent[c];
            if (r == (uint32_t)c) continue;
            for (;;) { const uint32_t q = *(volatile uint32_t *)&a.parent[r]; if (q == r) break; r = q; }
            a.parent[c] = r;
        }
        k4_grid_sync(a.ctl, epoch);
        for (unsigned long long v = gtid; v < n; v += gsize) a.comp[v] = (int32_t)a.parent[a.comp[v]];
        k4_grid_sync(a.ctl, epoch);
    }
}

__global__ void k4_split_kernel(const ulonglong2 *f, uint32_t nf, unsigned long long *key, unsigned long long *id)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf) { const ulonglong2 v = f[i]; key[i] = v.x; id[i] = v.y; }
}

template <typename T>
int k4_grow(T **p, uint64_t *cap, uint64_t need)
{
    if (*p && need <= *cap) return 0;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    TC_CUDA(cudaMalloc(p, std::max<uint64_t>(need, 1) * sizeof(T)));
    *cap = need;
    return 0;
}

}  // namespace

size_t tc_k4_xbuf_words(uint32_t n) { return (size_t)4 * n + 16; }   // 4 arrays + 8 epoch flags + 8 overflow notes

void tc_k4_free(tc_ctx *ctx)
{
    tc_k4_state &k = ctx->k4;
    for (int g = 0; g < 8; g++)
        if (k.peer_opened[g]) { cudaIpcCloseMemHandle(k.peer_xbuf[g]); k.peer_opened[g] = false; }
    cudaFree(k.d_xbuf); cudaFree(k.d_red); cudaFree(k.d_parent); cudaFree(k.d_out); cudaFree(k.d_fin); cudaFree(k.d_live); cudaFree(k.d_ctl); cudaFree(k.d_sort_tmp);
    if (k.h_out) cudaFreeHost(k.h_out);
    for (int e = 0; e < 3; e++) if (k.ev[e]) cudaEventDestroy(k.ev[e]);
    k = tc_k4_state();
}

// buffers for n vertices (grow-only).  The exchange buffer of a multi-GPU group is fixed when the group is formed.
int tc_k4_reserve(tc_ctx *ctx, uint32_t n)
{
    tc_k4_state &k = ctx->k4;
    if (k.nranks > 1) {
        TC_CHECK(n <= k.xbuf_n, "tc_mst: the multi-GPU group was formed for up to %u sketches, the set has %u", k.xbuf_n, n);
    } else if (tc_k4_xbuf_words(n) > k.cap_xbuf) {
        if (k4_grow(&k.d_xbuf, &k.cap_xbuf, tc_k4_xbuf_words(n))) return 1;
        k.xbuf_n = n;
    }
    if (k4_grow(&k.d_red, &k.cap_red, (uint64_t)2 * n) || k4_grow(&k.d_parent, &k.cap_parent, n) ||
        k4_grow(&k.d_out, &k.cap_out, (uint64_t)4 * n)) return 1;
    if (!k.d_ctl) TC_CUDA(cudaMalloc(&k.d_ctl, 16 * sizeof(uint32_t)));
    if ((uint64_t)n * 16 + 64 > k.h_out_bytes) {
        if (k.h_out) cudaFreeHost(k.h_out);
        k.h_out = nullptr; k.h_out_bytes = 0;
        TC_CUDA(cudaMallocHost(&k.h_out, (size_t)n * 16 + 64));
        k.h_out_bytes = (uint64_t)n * 16 + 64;
    }
    if (n + 1 > ctx->cap_comp) {
        cudaFree(ctx->d_comp); ctx->d_comp = nullptr; ctx->cap_comp = 0;
        TC_CUDA(cudaMalloc(&ctx->d_comp, (size_t)(n + 1) * 4));
        ctx->cap_comp = n + 1;
    }
    return 0;
}

// Queue the whole Boruvka over the candidate edges currently on the device (count read on the device: the pass need
// not have completed on the host) plus `nfin` explicit (key, id) edges, then the sort of the forest by (key, i, j).
// Nothing here waits for the device.
int tc_k4_enqueue(tc_ctx *ctx, int mode, const unsigned long long *d_fin_key, const unsigned long long *d_fin_id, uint32_t nfin)
{
    tc_k4_state &k = ctx->k4;
    const uint32_t n = ctx->n;
    cudaStream_t st = ctx->stream;
    if (tc_k4_reserve(ctx, n)) return 1;
    if (!k.grid) {
        int occ = 0;
        TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k4_boruvka_kernel, K4_THREADS, 0));
        TC_CHECK(occ >= 1, "tc_k4_enqueue: the Boruvka kernel does not fit an SM");
        int per_sm = std::min(occ, 2);
        if (const char *e = getenv("TCLUST_K4_CTAS_PER_SM")) per_sm = std::max(1, std::min(occ, atoi(e)));   // 1: two group members can share one GPU (tests)
        k.grid = ctx->num_sms * per_sm;
    }
    size_t tmp_bytes = 0;
    ulonglong2 *forest = reinterpret_cast<ulonglong2 *>(k.d_out);
    TC_CUDA(cub::DeviceMergeSort::SortKeys(nullptr, tmp_bytes, forest, (int)n, K4Less{}, st));
    if (k4_grow(&k.d_sort_tmp, &k.cap_sort_tmp, tmp_bytes + 256)) return 1;
    TC_CUDA(cudaMemsetAsync(k.d_out, 0xFF, (size_t)n * 16, st));
    TC_CUDA(cudaMemsetAsync(k.d_ctl, 0, 16 * sizeof(uint32_t), st));
    K4Args a{};
    a.edges = ctx->d_edges; a.edge_counter = ctx->d_edge_count; a.edge_cap = ctx->edge_cap;
    a.fin_key = d_fin_key; a.fin_id = d_fin_id; a.nfin = nfin;
    a.size = ctx->d_size; a.n = n; a.mode = mode; a.semb = ctx->edges_semb_size;
    a.comp = ctx->d_comp; a.parent = k.d_parent;
    a.rank = k.nranks > 1 ? k.rank : 0; a.nranks = std::max(1, k.nranks);
    for (int g = 0; g < 8; g++) a.xbuf[g] = nullptr;
    if (a.nranks > 1) { for (int g = 0; g < a.nranks; g++) a.xbuf[g] = g == a.rank ? k.d_xbuf : k.peer_xbuf[g]; }
    else a.xbuf[0] = k.d_xbuf;
    a.red = k.d_red;
    a.out = forest; a.ctl = k.d_ctl;
    {   // liveness flags of the edge chunks (written in round 0 before they are read)
        const uint64_t chunks = (ctx->edge_cap + K4_CHUNK - 1) / K4_CHUNK + 1;
        if (k4_grow(&k.d_live, &k.cap_live, chunks)) return 1;
        a.live = k.d_live;
    }
    a.xstride = k.xbuf_n;
    if (a.nranks > 1) {
        // The epoch flags are zeroed once, when the group is formed, and only grow: every rank launches the same
        // sequence of kernels, so the host-side base (advanced per launch by more barriers than a launch can use) is
        // the same on all of them
        a.sys_base = k.sys_base;
        k.sys_base += 1ull << 20;
    }
    for (int e = 0; e < 3; e++) if (!k.ev[e]) TC_CUDA(cudaEventCreate(&k.ev[e]));
    TC_CUDA(cudaEventRecord(k.ev[0], st));
    void *params[] = {(void *)&a};
    TC_CUDA(cudaLaunchCooperativeKernel((const void *)k4_boruvka_kernel, dim3(k.grid), dim3(K4_THREADS), params, 0, st));
    TC_CUDA(cudaEventRecord(k.ev[1], st));
    ctx->launches++;
    // forest sorted by (key, i, j) in one merge sort of the 16-byte entries (round 2, first version: two radix sorts of the
    // split arrays, 15 passes of ~13 us each for 10^5 entries); the unused tail (all ones) stays at the end
    size_t tb = tmp_bytes;
    TC_CUDA(cub::DeviceMergeSort::SortKeys(k.d_sort_tmp, tb, forest, (int)n, K4Less{}, st));
    TC_CUDA(cudaEventRecord(k.ev[2], st));
    ctx->launches += 2;
    return 0;
}

// Wait for the queued work and turn the forest into the reference's edges: {i, j<i, dist} sorted by (dist, i, j), dist by
// tc_dist_from_key (glibc log, the reference's expression order).  *n_edges_dev / *overflow: the candidate-edge counters
// of the pass that fed it.  comp_out (may be NULL): final component label of every vertex.
// One wait: control words, counters and the (at most n-1, padded to n) sorted forest entries come back together.
// direct != nullptr: the forest is written there (room for n edges) instead of into `mst`; *n_direct = its length
int tc_k4_collect(tc_ctx *ctx, int mode, int kmer_size, std::vector<tc_edge> &mst, uint64_t *n_edges_dev, bool *overflow,
                  std::vector<int32_t> *comp_out, tc_edge *direct, uint64_t *n_direct)
{
    tc_k4_state &k = ctx->k4;
    const uint32_t n = ctx->n;
    cudaStream_t st = ctx->stream;
    uint8_t *h = static_cast<uint8_t *>(k.h_out);
    uint32_t *h_ctl = reinterpret_cast<uint32_t *>(h);
    unsigned long long *h_cnt = reinterpret_cast<unsigned long long *>(h + 64 - 16);
    const ulonglong2 *h_f = reinterpret_cast<const ulonglong2 *>(h + 64);
    TC_CUDA(cudaMemcpyAsync(h_ctl, k.d_ctl, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TC_CUDA(cudaMemcpyAsync(h_cnt, ctx->d_edge_count, 16, cudaMemcpyDeviceToHost, st));
    TC_CUDA(cudaMemcpyAsync(h + 64, k.d_out, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    if (comp_out) {
        comp_out->resize(n);
        TC_CUDA(cudaMemcpyAsync(comp_out->data(), ctx->d_comp, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    }
    TC_CUDA(cudaStreamSynchronize(st));
    TC_CHECK(h_ctl[4] == 0, "tc_mst: the Boruvka kernel gave up at a barrier (%s)", h_ctl[4] == 2 ? "a peer GPU of the group did not arrive" : "grid barrier");
    const uint32_t nf = std::min(h_ctl[0], direct && n ? n - 1 : n);      // a forest has at most n - 1 edges (the caller's array)
    if (n_edges_dev) *n_edges_dev = h_cnt[0];
    if (overflow) *overflow = h_cnt[1] != 0 || h_cnt[0] > ctx->edge_cap || h_ctl[6] != 0;   // this rank's or (group) any rank's
    ctx->k4.last_rounds = h_ctl[5];
    if (direct) { mst.clear(); if (n_direct) *n_direct = nf; }
    else mst.resize(nf);
    tc_edge *dst = direct ? direct : mst.data();
    auto by_weight = [](const tc_edge &x, const tc_edge &y) {
        if (x.dist != y.dist) return x.dist < y.dist;
        if (x.pre != y.pre) return x.pre < y.pre;
        return x.suf < y.suf;
    };
    // key order is distance order, except where two different quotients round to the same distance: such a run is put
    // into (i, j) order (never seen on real keys; cheap to keep exact).  One distance per distinct key.
    double d = 0.0, prev_d = -1.0;
    uint32_t run0 = 0;                 // first edge of the current run of equal distances
    bool mixed = false, monotone = true;
    for (uint32_t e = 0; e < nf; e++) {
        if (e == 0 || h_f[e].x != h_f[e - 1].x) {
            d = tc_dist_from_key(h_f[e].x, mode, kmer_size);
            if (e > 0 && d == prev_d) mixed = true;
            else {
                if (mixed) std::sort(dst + run0, dst + e, by_weight);
                if (e > 0 && d < prev_d) monotone = false;
                run0 = e; mixed = false;
            }
            prev_d = d;
        }
        dst[e] = tc_edge{(int32_t)(h_f[e].y >> 32), (int32_t)(h_f[e].y & 0xFFFFFFFFull), d};
    }
    if (mixed) std::sort(dst + run0, dst + nf, by_weight);
    if (!monotone) std::sort(dst, dst + nf, by_weight);            // distances not monotone in the key (never for the valid modes)
    return 0;
}

// Bounded-memory MST: wait for the queued Boruvka and keep its (sorted) forest on the device as the explicit (key, id)
// edges of the next launch -- MSF(E1 u E2) = MSF(MSF(E1) u E2), what the reference does per task of 8 rows
// (src/MST.cpp:1373-1386, 571-572).  The candidate-edge counter restarts at 0.
int tc_k4_fold(tc_ctx *ctx, uint32_t *nfin)
{
    tc_k4_state &k = ctx->k4;
    const uint32_t n = ctx->n;
    cudaStream_t st = ctx->stream;
    uint32_t h_ctl[8];
    TC_CUDA(cudaMemcpyAsync(h_ctl, k.d_ctl, sizeof h_ctl, cudaMemcpyDeviceToHost, st));
    TC_CUDA(cudaStreamSynchronize(st));
    TC_CHECK(h_ctl[4] == 0, "tc_mst: the Boruvka kernel gave up at a barrier");
    TC_CHECK(h_ctl[6] == 0, "tc_mst: a batch of candidate edges overflowed the edge buffer (bounded-memory MST)");
    const uint32_t nf = std::min(h_ctl[0], n);
    if (k4_grow(&k.d_fin, &k.cap_fin, (uint64_t)2 * n)) return 1;
    if (nf) {
        k4_split_kernel<<<(nf + 255) / 256, 256, 0, st>>>(reinterpret_cast<const ulonglong2 *>(k.d_out), nf, k.d_fin, k.d_fin + n);
        TC_CUDA(cudaGetLastError());
    }
    TC_CUDA(cudaMemsetAsync(ctx->d_edge_count, 0, 2 * sizeof(unsigned long long), st));
    *nfin = nf;
    return 0;
}

// out[0] dense-pass kernel ms, [1] Boruvka kernel ms, [2] forest sort ms (CUDA events on the context's stream, last tc_mst),
// [3] candidate edges of this rank, [4] Boruvka rounds, [5] flagged tiles, [6] tiles in the lower triangle,
// [7] 1 when the pass ran on the reordered copy
extern "C" int tc_mst_stats(const tc_ctx *ctx, double *out)
{
    TC_CHECK(ctx && out, "tc_mst_stats: bad arguments");
    const tc_k4_state &k = ctx->k4;
    float ms = 0.f;
    out[0] = tc_last_pass_kernel_ms(ctx);
    out[1] = (k.ev[0] && cudaEventElapsedTime(&ms, k.ev[0], k.ev[1]) == cudaSuccess) ? ms : -1.0;
    out[2] = (k.ev[1] && cudaEventElapsedTime(&ms, k.ev[1], k.ev[2]) == cudaSuccess) ? ms : -1.0;
    cudaGetLastError();
    out[3] = (double)ctx->n_edges;
    out[4] = (double)k.last_rounds;
    const tc_ctx *src = (ctx->use_shadow && ctx->shadow) ? ctx->shadow : ctx;
    out[5] = src->tile_words ? (double)src->tiles_flagged : -1.0;
    const double nrb = (ctx->n + 31) / 32;
    out[6] = nrb * (nrb + 1) / 2;
    out[7] = (ctx->use_shadow && ctx->shadow) ? 1.0 : 0.0;
    out[8] = (double)ctx->bounded_batches; out[9] = (double)ctx->bounded_folds; out[10] = out[11] = 0.0;
    return 0;
}

// ------------------------------------------------------------------------------------- multi-GPU group
// The G contexts of one box that build one MST together.  Each context exports its exchange buffer
// (tc_mg_export), the handles travel by whatever plumbing the host has (torch.distributed all_gather between the
// processes of a torchrun job, a plain array between the threads of one process) and tc_mg_join maps the peers:
// raw pointers + peer access inside one process, CUDA IPC between processes.
namespace {
struct MgHandle {
    cudaIpcMemHandle_t ipc;
    unsigned long long ptr;
    long long pid;
    int device;
    uint32_t n_max;
};
static_assert(sizeof(MgHandle) <= TC_MG_HANDLE_BYTES, "tc_mg handle size");
}  // namespace

extern "C" int tc_mg_export(tc_ctx *ctx, uint32_t n_max, void *handle_out)
{
    TC_CHECK(ctx && handle_out && n_max > 0, "tc_mg_export: bad arguments");
    TC_CUDA(cudaSetDevice(ctx->device));
    tc_k4_state &k = ctx->k4;
    TC_CHECK(k.nranks <= 1, "tc_mg_export: the context is in a group already (tc_mg_leave first)");
    cudaFree(k.d_xbuf); k.d_xbuf = nullptr; k.cap_xbuf = 0;
    TC_CUDA(cudaMalloc(&k.d_xbuf, tc_k4_xbuf_words(n_max) * 8));       // its own allocation: IPC maps whole allocations
    k.cap_xbuf = tc_k4_xbuf_words(n_max); k.xbuf_n = n_max;
    TC_CUDA(cudaMemset(k.d_xbuf, 0, tc_k4_xbuf_words(n_max) * 8));      // epoch flags start at 0, before any peer can signal
    MgHandle h;
    memset(&h, 0, sizeof h);
    TC_CUDA(cudaIpcGetMemHandle(&h.ipc, k.d_xbuf));
    h.ptr = (unsigned long long)(uintptr_t)k.d_xbuf; h.pid = (long long)getpid(); h.device = ctx->device; h.n_max = n_max;
    memset(handle_out, 0, TC_MG_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof h);
    k.export_fresh = true;
    return 0;
}

extern "C" int tc_mg_join(tc_ctx *ctx, int rank, int nranks, const void *handles)
{
    TC_CHECK(ctx && handles && nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks, "tc_mg_join: bad rank %d of %d (at most 8 GPUs)", rank, nranks);
    TC_CUDA(cudaSetDevice(ctx->device));
    tc_k4_state &k = ctx->k4;
    const uint8_t *hb = static_cast<const uint8_t *>(handles);
    MgHandle me;
    memcpy(&me, hb + (size_t)rank * TC_MG_HANDLE_BYTES, sizeof me);
    TC_CHECK(me.ptr == (unsigned long long)(uintptr_t)k.d_xbuf && me.pid == (long long)getpid(), "tc_mg_join: handle %d is not this context's (tc_mg_export)", rank);
    TC_CHECK(k.export_fresh, "tc_mg_join: the exchange buffer has been used by an earlier group; call tc_mg_export again (its barrier flags must start at zero)");
    k.export_fresh = false;
    for (int g = 0; g < nranks; g++) {
        MgHandle h;
        memcpy(&h, hb + (size_t)g * TC_MG_HANDLE_BYTES, sizeof h);
        TC_CHECK(h.n_max == me.n_max, "tc_mg_join: rank %d exported for %u sketches, this rank for %u", g, h.n_max, me.n_max);
        if (g == rank) { k.peer_xbuf[g] = k.d_xbuf; continue; }
        if (h.pid == (long long)getpid()) {                     // same process: peer access + the raw pointer
            if (h.device != ctx->device) {                      // (two contexts on one device: tests of the group logic on a 1-GPU box)
                int can = 0;
                TC_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, h.device));
                TC_CHECK(can, "tc_mg_join: device %d cannot access device %d", ctx->device, h.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { tc_set_error("tc_mg_join: %s", cudaGetErrorString(e)); return 1; }
                cudaGetLastError();
            }
            k.peer_xbuf[g] = reinterpret_cast<unsigned long long *>((uintptr_t)h.ptr);
        } else {
            void *p = nullptr;
            TC_CUDA(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            k.peer_xbuf[g] = static_cast<unsigned long long *>(p);
            k.peer_opened[g] = true;
        }
    }
    k.rank = rank; k.nranks = nranks; k.sys_base = 0;
    return 0;
}

extern "C" int tc_mg_leave(tc_ctx *ctx)
{
    TC_CHECK(ctx, "tc_mg_leave: null context");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    tc_k4_state &k = ctx->k4;
    for (int g = 0; g < 8; g++) {
        if (k.peer_opened[g]) { cudaIpcCloseMemHandle(k.peer_xbuf[g]); k.peer_opened[g] = false; }
        k.peer_xbuf[g] = nullptr;
    }
    k.rank = 0; k.nranks = 0; k.sys_base = 0;
    return 0;
}

extern "C" int tc_mg_rank(const tc_ctx *ctx, int *rank, int *nranks)
{
    TC_CHECK(ctx, "tc_mg_rank: null context");
    if (rank) *rank = ctx->k4.nranks > 1 ? ctx->k4.rank : 0;
    if (nranks) *nranks = ctx->k4.nranks > 1 ? ctx->k4.nranks : 1;
    return 0;
}
