// tclust_api.cu -- C ABI (include/tclust.h): context, sketch residency (rank transform +
// group interleave), validation entry points, dense pass driver, Boruvka MST driver.
#include "tc_common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <cstring>
#include <numeric>

static thread_local std::string g_err;
void tc_set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
extern "C" const char *tc_last_error(void) { return g_err.c_str(); }
extern "C" int tc_version(void) { return TC_VERSION; }

int tc_scratch(tc_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->scratch_bytes) {
        if (ctx->d_scratch) cudaFree(ctx->d_scratch);
        ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
        TC_CUDA(cudaMalloc(&ctx->d_scratch, bytes));
        ctx->scratch_bytes = bytes;
    }
    *out = ctx->d_scratch;
    return 0;
}

extern "C" void tc_ctx_destroy(tc_ctx *ctx);
extern "C" int tc_ctx_create(int device, tc_ctx **out)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        tc_set_error("tc_ctx_create: no CUDA device (%s); libtclust_cuda has no CPU fallback",
                     cudaGetErrorString(e));
        return 1;
    }
    TC_CHECK(device >= 0 && device < ndev, "tc_ctx_create: device %d out of range (%d devices)", device, ndev);
    TC_CUDA(cudaSetDevice(device));
    if (const char *m = getenv("TCLUST_SYNC")) {          // how the host waits in cudaStreamSynchronize, for THIS device (set before its context is used)
        const unsigned f = !strcmp(m, "spin") ? cudaDeviceScheduleSpin : !strcmp(m, "yield") ? cudaDeviceScheduleYield
                         : !strcmp(m, "block") ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto;
        if (cudaSetDeviceFlags(f) != cudaSuccess) cudaGetLastError();
    }
    cudaDeviceProp prop;
    TC_CUDA(cudaGetDeviceProperties(&prop, device));
    TC_CHECK(prop.major >= 10, "tc_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
             device, prop.major, prop.minor);
    tc_ctx *ctx = new tc_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_pass0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_pass1);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_edge_count, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_edge_count, 0, 4 * sizeof(unsigned long long));
    if (e != cudaSuccess) {                               // a partly built context is torn down, not leaked
        tc_set_error("tc_ctx_create: %s", cudaGetErrorString(e));
        cudaGetLastError();
        tc_ctx_destroy(ctx);
        return 1;
    }
    *out = ctx;
    return 0;
}

// the big device buffers are kept between uploads (cudaFree/cudaMalloc of GB-sized buffers costs milliseconds, and
// tens of milliseconds once the context also holds the sort workspace): an upload only grows them
static void release_sketches(tc_ctx *ctx)
{
    cudaFree(ctx->d_hashes); cudaFree(ctx->d_off); cudaFree(ctx->d_size); cudaFree(ctx->d_ranks);
    cudaFree(ctx->d_gt); cudaFree(ctx->d_goff); cudaFree(ctx->d_edges); cudaFree(ctx->d_comp);
    cudaFree(ctx->d_kh_pivots); cudaFree(ctx->d_kh_pivoff);
    cudaFree(ctx->d_bu_keys); cudaFree(ctx->d_bu_mask); cudaFree(ctx->d_bu_count); cudaFree(ctx->d_multi);
    cudaFree(ctx->d_tile_flags); cudaFree(ctx->d_tile_rowcount);
    ctx->d_tile_flags = nullptr; ctx->d_tile_rowcount = nullptr; ctx->cap_tile_flags = ctx->cap_tile_rowcount = 0; ctx->tile_words = 0;
    ctx->d_multi = nullptr; ctx->cap_multi = 0;
    ctx->d_bu_keys = nullptr; ctx->d_bu_mask = nullptr; ctx->d_bu_count = nullptr; ctx->cap_bu = 0; ctx->cap_bu_count = 0;
    ctx->d_kh_pivots = nullptr; ctx->d_kh_pivoff = nullptr; ctx->cap_pivoff = 0; ctx->cap_pivots = 0;
    ctx->d_hashes = nullptr; ctx->d_off = nullptr; ctx->d_size = nullptr; ctx->d_ranks = nullptr;
    ctx->d_gt = nullptr; ctx->d_goff = nullptr; ctx->d_edges = nullptr; ctx->d_comp = nullptr;
    ctx->cap_hashes = ctx->cap_ranks = ctx->cap_gt = ctx->cap_n = ctx->cap_groups = ctx->cap_comp = 0;
    ctx->edge_cap = 0; ctx->edge_cap_for_n = 0;
}

static void reset_sketches(tc_ctx *ctx)
{
    ctx->kh_edge_units.clear(); ctx->kh_edge_units_n = 0; ctx->kh_units_on_device = false; ctx->kj_split_valid = false;
    ctx->n_edges = 0; ctx->n = 0; ctx->total = 0; ctx->ngroups = 0; ctx->edges_semb_size = 0;
    ctx->edge_cap_for_n = 0; ctx->use_shadow = false; ctx->pairs_cache_valid = false; ctx->gt_valid = false; ctx->edges_partial = false;
    tc_rep_table_reset(ctx);
}

template <typename T>
static int ensure_cap(T **p, uint64_t *cap, uint64_t need)
{
    if (*p && need <= *cap) return 0;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    TC_CUDA(cudaMalloc(p, std::max<uint64_t>(need, 1) * sizeof(T)));
    *cap = need;
    return 0;
}

extern "C" void tc_ctx_destroy(tc_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    reset_sketches(ctx);
    release_sketches(ctx);
    tc_shadow_free(ctx);
    tc_rep_table_free(ctx);
    tc_k1_free(ctx);
    tc_k2_free(ctx);
    tc_k4_free(ctx);
    cudaFree(ctx->d_edge_count);
    cudaFree(ctx->d_uf);
    cudaFree(ctx->d_scratch);
    for (int b = 0; b < 2; b++) {
        if (ctx->h_stage[b]) cudaFreeHost(ctx->h_stage[b]);
        if (ctx->ev_stage[b]) cudaEventDestroy(ctx->ev_stage[b]);
    }
    for (cudaEvent_t &e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    if (ctx->up_copy_stream) cudaStreamDestroy(ctx->up_copy_stream);
    cudaFree(ctx->d_kssd_dim); cudaFree(ctx->d_kssd_filter);
    cudaFree(ctx->d_kh_state); cudaFree(ctx->d_kh_units); cudaFree(ctx->d_kj_units); cudaFree(ctx->d_ks_tiles); cudaFree(ctx->d_ks_count);
    if (ctx->ev_pass0) cudaEventDestroy(ctx->ev_pass0);
    if (ctx->ev_pass1) cudaEventDestroy(ctx->ev_pass1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" uint64_t tc_kernel_launches(const tc_ctx *ctx) { return ctx->launches; }
extern "C" double tc_last_pass_kernel_ms(const tc_ctx *ctx)
{
    float ms = 0.f;
    if (!ctx->ev_pass0 || cudaEventElapsedTime(&ms, ctx->ev_pass0, ctx->ev_pass1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    return (double)ms;
}
extern "C" void *tc_stream(const tc_ctx *ctx) { return (void *)ctx->stream; }

// ------------------------------------------------------------------ rank transform kernels
__global__ void iota_kernel(uint32_t *p, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void iota_base_kernel(uint32_t *p, uint64_t n, uint32_t base)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = base + (uint32_t)i;
}
__global__ void head_flag_kernel(const uint64_t *keys, uint32_t *flag, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void scatter_rank_kernel(const uint32_t *scan, const uint32_t *perm, uint32_t *ranks, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ranks[perm[i]] = scan[i] - 1u;  // inclusive scan of head flags, 0-based dense rank
}
struct HeadFlagOp {
    const uint64_t *keys;
    __host__ __device__ uint32_t operator()(uint64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; }
};
// scatter_rank_kernel + multi_bitmap_kernel in one pass over the scan
__global__ void scatter_rank_multi_kernel(const uint32_t *scan, const uint32_t *perm, uint32_t *ranks, uint64_t n, uint32_t *multi)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = scan[i];
    ranks[perm[i]] = v - 1u;
    if (i + 1 < n && scan[i + 1] == v && (i == 0 || scan[i - 1] != v)) atomicOr(&multi[(v - 1u) >> 5], 1u << ((v - 1u) & 31));
}
// scan = inclusive scan of the head flags over the sorted hashes: rank r sits at the positions holding r + 1; the first
// position of a run of two or more marks its rank in the bitmap
__global__ void multi_bitmap_kernel(const uint32_t *scan, uint64_t n, uint32_t *multi)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const uint32_t v = scan[i];
    if (scan[i + 1] == v && (i == 0 || scan[i - 1] != v)) atomicOr(&multi[(v - 1u) >> 5], 1u << ((v - 1u) & 31));
}
// every sketch must be strictly ascending (sorted, duplicate-free): the merge and the hash join rely on it
__global__ void check_sorted_kernel(const uint64_t *h, const unsigned long long *off, uint32_t n, unsigned int *bad)
{
    const uint32_t g = blockIdx.x;
    if (g >= n) return;
    const unsigned long long o = off[g], len = off[g + 1] - o;
    for (unsigned long long i = threadIdx.x + 1; i < len; i += blockDim.x)
        if (h[o + i] <= h[o + i - 1]) { atomicMin(bad, g); return; }
}
// the same check for sketches [g0, g0 + gridDim.x): upload_sorted_runs checks a sketch as soon as its last hash has landed
__global__ void check_sorted_range_kernel(const uint64_t *h, const unsigned long long *off, uint32_t g0, unsigned int *bad)
{
    const uint32_t g = g0 + blockIdx.x;
    const unsigned long long o = off[g], len = off[g + 1] - o;
    for (unsigned long long i = threadIdx.x + 1; i < len; i += blockDim.x)
        if (h[o + i] <= h[o + i - 1]) { atomicMin(bad, g); return; }
}
__global__ void sizes_kernel(const unsigned long long *off, uint32_t *size, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) size[i] = (uint32_t)(off[i + 1] - off[i]);
}
// gt[goff[g] + row*32 + lane] = rank `row` of sketch g*32+lane (TC_SENT_A padding)
__global__ void interleave_kernel(const uint32_t *ranks, const unsigned long long *off, const unsigned long long *goff,
                                  uint32_t n, uint32_t ngroups, uint32_t *gt)
{
    const uint32_t g = blockIdx.y;
    const unsigned long long rows = (goff[g + 1] - goff[g]) / TC_GROUP;
    const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * TC_GROUP) return;
    const uint32_t lane = (uint32_t)(e & 31), row = (uint32_t)(e >> 5);
    const uint32_t sk = g * TC_GROUP + lane;
    uint32_t v = TC_SENT_A;
    if (sk < n) {
        const unsigned long long o = off[sk], len = off[sk + 1] - o;
        if (row < len) v = ranks[o + row];
    }
    gt[goff[g] + e] = v;
}

static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

// TCLUST_TIMING=1: phase times of the upload on stderr (development aid; synchronises the stream at every mark)
struct PhaseTimer {
    bool on; cudaStream_t st; std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(cudaStream_t s) : on(getenv("TCLUST_TIMING") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what)
    {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[tclust timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---- sort workspace of the rank transform: two key and two value buffers + cub's temporary storage, carved out of the
// context's scratch (24 B per hash + temporary storage; kept between uploads)
constexpr int UP_CHUNKS = TC_UP_CHUNKS;
struct UploadSort {
    uint64_t *k_a = nullptr, *k_b = nullptr;
    uint32_t *v_a = nullptr, *v_b = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    const uint64_t *keys = nullptr;      // the globally sorted hashes ...
    const uint32_t *vals = nullptr;      // ... and the input position of each
    uint32_t *spare = nullptr;           // the value buffer not holding `vals`
};

static int upload_sort_workspace(tc_ctx *ctx, uint64_t total, uint64_t clen, UploadSort *us)
{
    cudaStream_t st = ctx->stream;
    size_t tmp_bytes = 0, t = 0;
    TC_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)total, st));
    if (clen == 0) {
        cub::DoubleBuffer<uint64_t> kb0(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t> vb0(nullptr, nullptr);
        TC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t, kb0, vb0, (int)total, 0, 64, st));
        tmp_bytes = std::max(tmp_bytes, t);
    } else {
        // out-of-place sort of one piece (cub keeps its alternate buffers in the temporary storage) and the merges
        TC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                                (uint32_t *)nullptr, (int)clen, 0, 64, st));
        tmp_bytes = std::max(tmp_bytes, t);
        for (uint64_t w = clen; w < total; w *= 2) {
            const uint64_t a = std::min(w, total), b = std::min(w, total - a);
            t = 0;
            TC_CUDA(cub::DeviceMerge::MergePairs(nullptr, t, (const uint64_t *)nullptr, (const uint32_t *)nullptr, (int)a,
                                                 (const uint64_t *)nullptr, (const uint32_t *)nullptr, (int)b, (uint64_t *)nullptr,
                                                 (uint32_t *)nullptr, ::cuda::std::less<>{}, st));
            tmp_bytes = std::max(tmp_bytes, t);
        }
    }
    tmp_bytes = (tmp_bytes + 255) & ~(size_t)255;
    const size_t a256 = 255;
    const size_t o_kb = (total * 8 + a256) & ~a256, o_va = o_kb + ((total * 8 + a256) & ~a256);
    const size_t o_vb = o_va + ((total * 4 + a256) & ~a256), o_tmp = o_vb + ((total * 4 + a256) & ~a256);
    void *ws = nullptr;
    if (tc_scratch(ctx, o_tmp + tmp_bytes, &ws)) return 1;
    uint8_t *wsb = static_cast<uint8_t *>(ws);
    us->k_a = reinterpret_cast<uint64_t *>(wsb); us->k_b = reinterpret_cast<uint64_t *>(wsb + o_kb);
    us->v_a = reinterpret_cast<uint32_t *>(wsb + o_va); us->v_b = reinterpret_cast<uint32_t *>(wsb + o_vb);
    us->tmp = wsb + o_tmp; us->tmp_bytes = tmp_bytes;
    return 0;
}

// Host hashes -> d_hashes in UP_CHUNKS pieces on the copy stream; on the work stream piece c is sorted (with its input
// positions as values) into run c of buffer B as soon as its copy has landed, and whenever two neighbouring runs of one
// width are complete they are merged into the other buffer (binary-counter order), so the work left when the last
// byte arrives is one piece's sort and log2(UP_CHUNKS) merges instead of the sort of the whole set.
// The input contract (every sketch strictly ascending) is checked piece by piece as well: sketch g is checked by the
// piece that brings its last hash (`off`: host offsets, `d_bad`: smallest offending sketch, read by the caller later), so
// that only the last piece's sketches are checked after the copy.
static int upload_sorted_runs(tc_ctx *ctx, const uint64_t *hashes, uint64_t total, UploadSort *us, const uint64_t *off, uint32_t n,
                              unsigned int *d_bad)
{
    cudaStream_t st = ctx->stream;
    if (!ctx->up_copy_stream) TC_CUDA(cudaStreamCreateWithFlags(&ctx->up_copy_stream, cudaStreamNonBlocking));
    for (int c = 0; c <= UP_CHUNKS; c++)
        if (!ctx->ev_chunk[c]) TC_CUDA(cudaEventCreateWithFlags(&ctx->ev_chunk[c], cudaEventDisableTiming));
    const uint64_t clen = ((total + UP_CHUNKS - 1) / UP_CHUNKS + 1023) & ~1023ull;
    if (upload_sort_workspace(ctx, total, clen, us)) return 1;
    // the copies start once everything queued on the work stream so far (a previous set's kernels) is done
    TC_CUDA(cudaEventRecord(ctx->ev_chunk[UP_CHUNKS], st));
    TC_CUDA(cudaStreamWaitEvent(ctx->up_copy_stream, ctx->ev_chunk[UP_CHUNKS], 0));
    iota_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(us->v_a, total);
    // level 0 (sorted pieces) lives in B, level 1 in A, level 2 in B, ...
    auto kbuf = [&](int level) { return (level & 1) ? us->k_a : us->k_b; };
    auto vbuf = [&](int level) { return (level & 1) ? us->v_a : us->v_b; };
    int top = 0;
    while ((1 << top) < UP_CHUNKS) top++;
    // Piece c's copy is queued, then its sort and merges, then piece c+1's copy: from page-locked memory all of it is
    // asynchronous and the copy queue never runs dry (a piece takes ~2 ms, queueing its sort ~0.1 ms); from pageable memory
    // (a std::vector of the reference) cudaMemcpyAsync returns when the piece is staged, and the work queued before it runs
    // on the device in the meantime.
    for (int c = 0; c < UP_CHUNKS; c++) {
        const uint64_t pos = std::min(total, (uint64_t)c * clen), len = std::min(clen, total - pos);
        if (len) TC_CUDA(cudaMemcpyAsync(ctx->d_hashes + pos, hashes + pos, len * 8, cudaMemcpyHostToDevice, ctx->up_copy_stream));
        TC_CUDA(cudaEventRecord(ctx->ev_chunk[c], ctx->up_copy_stream));
        TC_CUDA(cudaStreamWaitEvent(st, ctx->ev_chunk[c], 0));
        if (len) {
            // sketches whose last hash lies in (pos, pos + len]: off[g + 1] in that range
            const uint32_t g_lo = (uint32_t)(std::upper_bound(off + 1, off + n + 1, pos) - (off + 1));
            const uint32_t g_hi = (uint32_t)(std::upper_bound(off + 1, off + n + 1, pos + len) - (off + 1));
            if (g_hi > g_lo) {
                check_sorted_range_kernel<<<g_hi - g_lo, 128, 0, st>>>(ctx->d_hashes, reinterpret_cast<const unsigned long long *>(ctx->d_off),
                                                                        g_lo, d_bad);
                ctx->launches++;
            }
        }
        if (len) {
            size_t tb = us->tmp_bytes;
            TC_CUDA(cub::DeviceRadixSort::SortPairs(us->tmp, tb, (const uint64_t *)(ctx->d_hashes + pos), us->k_b + pos,
                                                    (const uint32_t *)(us->v_a + pos), us->v_b + pos, (int)len, 0, 64, st));
        }
        for (int level = 1; level <= top && ((c + 1) & ((1 << level) - 1)) == 0; level++) {
            const uint64_t w = clen << (level - 1);                        // width of the two runs being merged
            const uint64_t p0 = std::min(total, (uint64_t)(c + 1 - (1 << level)) * clen);
            const uint64_t la = std::min(w, total - p0), lb = std::min(w, total - p0 - la);
            if (la + lb == 0) continue;
            if (lb == 0) {
                TC_CUDA(cudaMemcpyAsync(kbuf(level) + p0, kbuf(level - 1) + p0, la * 8, cudaMemcpyDeviceToDevice, st));
                TC_CUDA(cudaMemcpyAsync(vbuf(level) + p0, vbuf(level - 1) + p0, la * 4, cudaMemcpyDeviceToDevice, st));
                continue;
            }
            size_t tb = us->tmp_bytes;
            TC_CUDA(cub::DeviceMerge::MergePairs(us->tmp, tb, (const uint64_t *)(kbuf(level - 1) + p0),
                                                 (const uint32_t *)(vbuf(level - 1) + p0), (int)la,
                                                 (const uint64_t *)(kbuf(level - 1) + p0 + la),
                                                 (const uint32_t *)(vbuf(level - 1) + p0 + la), (int)lb, kbuf(level) + p0,
                                                 vbuf(level) + p0, ::cuda::std::less<>{}, st));
        }
    }
    us->keys = kbuf(top); us->vals = vbuf(top); us->spare = vbuf(top + 1);
    ctx->launches += 2 * UP_CHUNKS;
    return 0;
}
// (Tried in round 2: four or five pieces of decreasing size merged into an accumulating prefix, so that only one merge is
// left at the end -- 23.3 / 22.8 ms per upload against 22.4 ms for this scheme: the radix sort runs only 1.5x faster than
// the PCIe copy delivers, so the sorts of the larger early pieces pile up and the tail does not shrink.)

// gt[goff[g] + row*32 + lane] layout of the resident ranks for k3_tile.cu, built when a merge-kernel entry point needs it
int tc_ensure_interleaved(tc_ctx *ctx)
{
    if (ctx->gt_valid) return 0;
    const uint32_t n = ctx->n;
    cudaStream_t st = ctx->stream;
    std::vector<uint64_t> goff(ctx->ngroups + 1, 0);
    uint64_t max_rows = 0;
    for (uint32_t g = 0; g < ctx->ngroups; g++) {
        uint32_t rows = 0;
        for (uint32_t l = 0; l < TC_GROUP && g * TC_GROUP + l < n; l++) rows = std::max(rows, ctx->h_size[g * TC_GROUP + l]);
        goff[g + 1] = goff[g] + (uint64_t)rows * TC_GROUP;
        max_rows = std::max<uint64_t>(max_rows, rows);
    }
    if (ensure_cap(&ctx->d_goff, &ctx->cap_groups, (uint64_t)ctx->ngroups + 1)) return 1;
    TC_CUDA(cudaMemcpyAsync(ctx->d_goff, goff.data(), (size_t)(ctx->ngroups + 1) * 8, cudaMemcpyHostToDevice, st));
    if (ensure_cap(&ctx->d_gt, &ctx->cap_gt, std::max<uint64_t>(goff[ctx->ngroups], 1))) return 1;
    if (max_rows > 0) {
        dim3 grid(nblk(max_rows * TC_GROUP, 256), ctx->ngroups);
        interleave_kernel<<<grid, 256, 0, st>>>(ctx->d_ranks, reinterpret_cast<unsigned long long *>(ctx->d_off),
                                                reinterpret_cast<unsigned long long *>(ctx->d_goff), n, ctx->ngroups,
                                                ctx->d_gt);
        ctx->launches++;
    }
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(st));          // `goff` (host) must outlive its copy
    ctx->gt_valid = true;
    return 0;
}

// the hashes arrive as `nruns` runs already sorted by hash (keys ascending inside run r = [run_off[r], run_off[r+1])),
// each with the position its hash has in the flat array (tc_sketches_upload_runs_dev)
struct SortedRuns { const uint64_t *keys; const uint32_t *pos; const uint64_t *run_off; uint32_t nruns; };

__global__ void check_sorted_ranks_kernel(const uint32_t *r, const unsigned long long *off, uint32_t n, unsigned int *bad)
{
    const uint32_t g = blockIdx.x;
    if (g >= n) return;
    const unsigned long long o = off[g], len = off[g + 1] - o;
    for (unsigned long long i = threadIdx.x + 1; i < len; i += blockDim.x)
        if (r[o + i] <= r[o + i - 1]) { atomicMin(bad, g); return; }
}

// merge the runs pairwise (level by level, ping-pong between the caller's buffers and the workspace) into one sorted run
static int merge_sorted_runs(tc_ctx *ctx, const SortedRuns &r, uint64_t total, UploadSort *us)
{
    cudaStream_t st = ctx->stream;
    uint64_t longest = 0;
    for (uint32_t i = 0; i < r.nruns; i++) longest = std::max(longest, r.run_off[i + 1] - r.run_off[i]);
    if (upload_sort_workspace(ctx, total, std::max<uint64_t>(longest, 1024), us)) return 1;
    std::vector<uint64_t> bounds(r.run_off, r.run_off + r.nruns + 1);
    const uint64_t *ks = r.keys; const uint32_t *vs = r.pos;
    uint64_t *kd = us->k_a; uint32_t *vd = us->v_a;
    int level = 0;
    if (bounds.size() == 2) {          // one run: it is the sorted array already; copy (the caller's buffer is not ours to keep)
        TC_CUDA(cudaMemcpyAsync(us->k_a, ks, total * 8, cudaMemcpyDeviceToDevice, st));
        TC_CUDA(cudaMemcpyAsync(us->v_a, vs, total * 4, cudaMemcpyDeviceToDevice, st));
        us->keys = us->k_a; us->vals = us->v_a; us->spare = us->v_b;
        return 0;
    }
    while (bounds.size() > 2) {
        std::vector<uint64_t> nb{bounds[0]};
        for (size_t i = 0; i + 1 < bounds.size(); i += 2) {
            const uint64_t a0 = bounds[i], a1 = bounds[i + 1], b1 = i + 2 < bounds.size() ? bounds[i + 2] : a1;
            if (b1 == a1) {
                if (a1 > a0) {
                    TC_CUDA(cudaMemcpyAsync(kd + a0, ks + a0, (a1 - a0) * 8, cudaMemcpyDeviceToDevice, st));
                    TC_CUDA(cudaMemcpyAsync(vd + a0, vs + a0, (a1 - a0) * 4, cudaMemcpyDeviceToDevice, st));
                }
            } else {
                size_t tb = us->tmp_bytes;
                TC_CUDA(cub::DeviceMerge::MergePairs(us->tmp, tb, ks + a0, vs + a0, (int)(a1 - a0), ks + a1, vs + a1, (int)(b1 - a1),
                                                     kd + a0, vd + a0, ::cuda::std::less<>{}, st));
                ctx->launches += 2;
            }
            nb.push_back(b1);
        }
        bounds.swap(nb);
        ks = kd; vs = vd;
        level++;
        kd = (level & 1) ? us->k_b : us->k_a; vd = (level & 1) ? us->v_b : us->v_a;
    }
    us->keys = ks; us->vals = vs;
    us->spare = const_cast<uint32_t *>(vs) == us->v_a ? us->v_b : us->v_a;
    return 0;
}

// ranks_flag: the rank transform is done already (value-range sharded upload): per element its global dense rank with bit 31
// set when the rank occurs in more than one sketch
__global__ void ranks_from_flagged_kernel(const uint32_t *rf, uint64_t n, uint32_t *ranks, uint32_t *multi)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = rf[i], r = v & 0x7FFFFFFFu;
    ranks[i] = r;
    if (v >> 31) atomicOr(&multi[r >> 5], 1u << (r & 31));
}

static int upload_common(tc_ctx *ctx, const uint64_t *hashes, bool hashes_on_device, const uint64_t *off, uint32_t n,
                         const SortedRuns *runs = nullptr, const uint32_t *ranks_flag = nullptr)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    PhaseTimer pt(ctx->stream);
    reset_sketches(ctx);
    pt.mark("reset previous set");
    TC_CHECK(n > 0, "tc_sketches_upload: n == 0");
    const uint64_t total = off[n];
    // (cub's merge and scan entry points used by the rank transform count items in an int)
    TC_CHECK(total <= 0x7FFFFFFFull, "tc_sketches_upload: %llu hashes exceed the 2^31 - 1 one context ranks at once",
             (unsigned long long)total);
    ctx->n = n; ctx->total = total; ctx->hashes_resident = true;
    ctx->h_off.assign(off, off + n + 1);
    ctx->h_size.resize(n);
    uint32_t max_size = 0;
    for (uint32_t i = 0; i < n; i++) {
        TC_CHECK(off[i + 1] >= off[i], "tc_sketches_upload: offsets not monotone at %u", i);
        ctx->h_size[i] = (uint32_t)(off[i + 1] - off[i]);
        max_size = std::max(max_size, ctx->h_size[i]);
    }
    cudaStream_t st = ctx->stream;
    const uint64_t tot1 = std::max<uint64_t>(total, 1);
    const bool no_hashes = runs || ranks_flag;      // the flat u64 hashes are not brought to this device
    if ((!no_hashes && ensure_cap(&ctx->d_hashes, &ctx->cap_hashes, tot1)) || ensure_cap(&ctx->d_ranks, &ctx->cap_ranks, tot1)) return 1;
    if (n + 1 > ctx->cap_n) {
        cudaFree(ctx->d_off); cudaFree(ctx->d_size);
        ctx->d_off = nullptr; ctx->d_size = nullptr; ctx->cap_n = 0;
        TC_CUDA(cudaMalloc(&ctx->d_off, (size_t)(n + 1) * 8));
        TC_CUDA(cudaMalloc(&ctx->d_size, (size_t)(n + 1) * 4));
        ctx->cap_n = n + 1;
    }
    pt.mark("host prep + cudaMalloc");
    // Host hashes of a large set arrive in UP_CHUNKS pieces on a copy stream; every piece is radix-sorted as soon as it
    // has landed and the sorted runs are merged pairwise (cub::DeviceMerge), so that all but the last piece's sort and
    // the last merges run under the PCIe copy (upload_sorted_runs below).  TCLUST_UPLOAD_OVERLAP=0: one copy, one sort; =1: pieces whatever the size.
    const char *ov = getenv("TCLUST_UPLOAD_OVERLAP");
    const bool ov_off = ov && !strcmp(ov, "0"), ov_force = ov && !strcmp(ov, "1");   // 1: whatever the size (tests)
    const bool chunked = !no_hashes && !hashes_on_device && hashes != ctx->d_hashes && !ov_off &&
                         (total >= (16ull << 20) || (ov_force && total >= 2));
    UploadSort us;
    unsigned int *d_bad_chunked = reinterpret_cast<unsigned int *>(ctx->d_edge_count + 3);
    if (no_hashes) {
        // sorted runs / finished ranks from a multi-GPU upload: no copy of the flat hashes (d_hashes stays empty: only the
        // independent validation kernel reads it), the order check runs on the ranks below
        ctx->hashes_resident = false;
    } else if (chunked) {
        TC_CUDA(cudaMemcpyAsync(ctx->d_off, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        TC_CUDA(cudaMemsetAsync(d_bad_chunked, 0xFF, 4, st));
        if (upload_sorted_runs(ctx, hashes, total, &us, off, n, d_bad_chunked)) return 1;
    } else if (hashes != ctx->d_hashes)   // tc_sketches_upload_rows has staged them in place already
        TC_CUDA(cudaMemcpyAsync(ctx->d_hashes, hashes, total * 8,
                                hashes_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!chunked) TC_CUDA(cudaMemcpyAsync(ctx->d_off, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    sizes_kernel<<<nblk(n, 256), 256, 0, st>>>(reinterpret_cast<unsigned long long *>(ctx->d_off), ctx->d_size, n);
    ctx->launches++;
    if (!no_hashes && !chunked) {   // input contract check (0.8 GB read at N = 100k: ~0.2 ms)
        unsigned int *d_bad = reinterpret_cast<unsigned int *>(ctx->d_edge_count + 3), h_bad = 0xFFFFFFFFu;
        TC_CUDA(cudaMemsetAsync(d_bad, 0xFF, 4, st));
        check_sorted_kernel<<<n, 128, 0, st>>>(ctx->d_hashes, reinterpret_cast<unsigned long long *>(ctx->d_off), n, d_bad);
        ctx->launches++;
        TC_CUDA(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
        TC_CUDA(cudaStreamSynchronize(st));
        if (h_bad != 0xFFFFFFFFu) {
            const uint32_t g = h_bad;
            reset_sketches(ctx);
            tc_set_error("tc_sketches_upload: sketch %u is not strictly ascending (sketches must be sorted and duplicate-free, "
                         "as Sketch::MinHash::storeMinHashes returns them)", g);
            return 1;
        }
    }
    pt.mark("copy of the hashes + check");

    // ---- order-preserving dense rank transform: u64 hash -> u32 rank (equal hash <=> equal rank)
    if (total == 0) {                      // nothing can be shared: an empty bitmap keeps the kernels' argument valid
        if (ensure_cap(&ctx->d_multi, &ctx->cap_multi, 1)) return 1;
        TC_CUDA(cudaMemsetAsync(ctx->d_multi, 0, 4, ctx->stream));
    }
    if (total > 0 && ranks_flag) {
        const uint64_t mwords = total / 32 + 1;
        if (ensure_cap(&ctx->d_multi, &ctx->cap_multi, mwords)) return 1;
        TC_CUDA(cudaMemsetAsync(ctx->d_multi, 0, mwords * 4, st));
        ranks_from_flagged_kernel<<<nblk(total, 256), 256, 0, st>>>(ranks_flag, total, ctx->d_ranks, ctx->d_multi);
        ctx->launches++;
        pt.mark("ranks + bitmap from the sharded transform");
    } else if (total > 0) {
        if (runs) {
            if (merge_sorted_runs(ctx, *runs, total, &us)) return 1;
        } else if (!chunked) {
            // sort workspace: kept in the context's scratch between uploads (24 B per hash + cub's temporary storage)
            if (upload_sort_workspace(ctx, total, 0, &us)) return 1;
            pt.mark("sort workspace");
            TC_CUDA(cudaMemcpyAsync(us.k_a, ctx->d_hashes, total * 8, cudaMemcpyDeviceToDevice, st));
            iota_kernel<<<nblk(total, 256), 256, 0, st>>>(us.v_a, total);
            cub::DoubleBuffer<uint64_t> kb(us.k_a, us.k_b);
            cub::DoubleBuffer<uint32_t> vb(us.v_a, us.v_b);
            TC_CUDA(cub::DeviceRadixSort::SortPairs(us.tmp, us.tmp_bytes, kb, vb, (int)total, 0, 64, st));
            us.keys = kb.Current(); us.vals = vb.Current(); us.spare = vb.Alternate();
        }
        pt.mark("radix sort");
        uint32_t *flag = us.spare;        // reuse the spare value buffer
        size_t tmp2 = us.tmp_bytes;
        // inclusive scan of the head flags, the flags computed on the fly from the sorted keys (no flag array is written and
        // read back), then one kernel scatters the ranks and marks the ranks that occur in more than one sketch
        // (sketches are duplicate-free, so a run of equal hashes spans sketches)
        cub::CountingInputIterator<uint64_t> counting(0);
        cub::TransformInputIterator<uint32_t, HeadFlagOp, cub::CountingInputIterator<uint64_t>> heads(counting, HeadFlagOp{us.keys});
        TC_CUDA(cub::DeviceScan::InclusiveSum(us.tmp, tmp2, heads, flag, (int)total, st));
        const uint64_t mwords = total / 32 + 1;
        if (ensure_cap(&ctx->d_multi, &ctx->cap_multi, mwords)) return 1;
        TC_CUDA(cudaMemsetAsync(ctx->d_multi, 0, mwords * 4, st));
        scatter_rank_multi_kernel<<<nblk(total, 256), 256, 0, st>>>(flag, us.vals, ctx->d_ranks, total, ctx->d_multi);
        ctx->launches += 12;  // iota, sort passes (cub), flags, scan, scatter (approximate; cub internals)
        unsigned int h_bad = 0xFFFFFFFFu;
        if (chunked) TC_CUDA(cudaMemcpyAsync(&h_bad, d_bad_chunked, 4, cudaMemcpyDeviceToHost, st));
        TC_CUDA(cudaStreamSynchronize(st));
        if (h_bad != 0xFFFFFFFFu) {         // verdict of the piece-wise contract check (upload_sorted_runs)
            const uint32_t g = h_bad;
            reset_sketches(ctx);
            tc_set_error("tc_sketches_upload: sketch %u is not strictly ascending (sketches must be sorted and duplicate-free, "
                         "as Sketch::MinHash::storeMinHashes returns them)", g);
            return 1;
        }
        pt.mark("flags + scan + scatter");
    }
    if (no_hashes && total > 0) {      // the contract check on the ranks (order-preserving: same verdict as on the hashes)
        unsigned int *d_bad = reinterpret_cast<unsigned int *>(ctx->d_edge_count + 3), h_bad = 0xFFFFFFFFu;
        TC_CUDA(cudaMemsetAsync(d_bad, 0xFF, 4, st));
        check_sorted_ranks_kernel<<<n, 128, 0, st>>>(ctx->d_ranks, reinterpret_cast<unsigned long long *>(ctx->d_off), n, d_bad);
        ctx->launches++;
        TC_CUDA(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
        TC_CUDA(cudaStreamSynchronize(st));
        if (h_bad != 0xFFFFFFFFu) {
            const uint32_t g = h_bad;
            reset_sketches(ctx);
            tc_set_error("tc_sketches_upload: sketch %u is not strictly ascending (sketches must be sorted and duplicate-free, "
                         "as Sketch::MinHash::storeMinHashes returns them)", g);
            return 1;
        }
    }

    // ---- group-interleaved layout: only the first-generation merge kernel (semantic B, validation) reads it; built
    //      on first use (tc_ensure_interleaved)
    ctx->ngroups = (n + TC_GROUP - 1) / TC_GROUP;
    ctx->gt_valid = false;
    int rc = tc_kh_prepare(ctx);
    pt.mark("distinct counts + pivots");
    if (!rc) rc = tc_shadow_build(ctx);
    pt.mark("reordered copy");
    return rc;
}

extern "C" int tc_sketches_upload(tc_ctx *ctx, const uint64_t *hashes, const uint64_t *off, uint32_t n)
{
    return upload_common(ctx, hashes, false, off, n);
}
extern "C" int tc_sketches_upload_dev(tc_ctx *ctx, const uint64_t *hashes_dev, const uint64_t *off, uint32_t n)
{
    return upload_common(ctx, hashes_dev, true, off, n);
}

// ---- multi-GPU upload: every rank sorts its share of the hashes, the sorted runs are all-gathered, every rank merges
extern "C" int tc_upload_sort_share_dev(tc_ctx *ctx, const uint64_t *hashes_dev, uint64_t len, uint64_t pos_base,
                                        uint64_t *keys_out_dev, uint32_t *pos_out_dev)
{
    TC_CHECK(ctx && (len == 0 || (hashes_dev && keys_out_dev && pos_out_dev)), "tc_upload_sort_share_dev: bad arguments");
    TC_CHECK(pos_base + len <= 0x7FFFFFFFull, "tc_upload_sort_share_dev: positions exceed 2^31 - 1");
    TC_CUDA(cudaSetDevice(ctx->device));
    if (!len) return 0;
    cudaStream_t st = ctx->stream;
    size_t tmp_bytes = 0;
    TC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                            (uint32_t *)nullptr, (int)len, 0, 64, st));
    void *ws = nullptr;
    const size_t o_tmp = (len * 4 + 255) & ~(size_t)255;
    if (tc_scratch(ctx, o_tmp + tmp_bytes, &ws)) return 1;
    uint32_t *iota = static_cast<uint32_t *>(ws);
    iota_base_kernel<<<nblk(len, 256), 256, 0, st>>>(iota, len, (uint32_t)pos_base);
    TC_CUDA(cub::DeviceRadixSort::SortPairs(static_cast<uint8_t *>(ws) + o_tmp, tmp_bytes, hashes_dev, keys_out_dev, (const uint32_t *)iota,
                                            pos_out_dev, (int)len, 0, 64, st));
    ctx->launches += 10;
    TC_CUDA(cudaStreamSynchronize(st));          // the outputs are handed to the caller's collective next
    return 0;
}

extern "C" int tc_sketches_upload_runs_dev(tc_ctx *ctx, const uint64_t *keys_dev, const uint32_t *pos_dev, const uint64_t *run_off,
                                           uint32_t nruns, const uint64_t *off, uint32_t n)
{
    TC_CHECK(ctx && keys_dev && pos_dev && run_off && off && nruns >= 1, "tc_sketches_upload_runs_dev: bad arguments");
    TC_CHECK(run_off[0] == 0 && run_off[nruns] == off[n], "tc_sketches_upload_runs_dev: the runs hold %llu hashes, the offsets %llu",
             (unsigned long long)run_off[nruns], (unsigned long long)off[n]);
    for (uint32_t r = 0; r < nruns; r++) TC_CHECK(run_off[r + 1] >= run_off[r], "tc_sketches_upload_runs_dev: run offsets not monotone");
    const SortedRuns runs{keys_dev, pos_dev, run_off, nruns};
    return upload_common(ctx, nullptr, true, off, n, &runs);
}

// ---- value-range sharded rank transform (multi-GPU upload, second form): no rank sorts or gathers more than it must.
// Pieces (the exchanges between them are the caller's: all-to-all / all-gather, plumbing):
//   tc_vr_partition_dev   pairs (a, b) -> grouped by bucket: mode 0 bucket = number of splitters <= a (a = hash, b = its
//                         position, generated as pos_base + index when b_dev is NULL); mode 1 bucket = b / per (a = rank).
//   tc_vr_local_ranks_dev sort the (hash, position) pairs of this rank's value range, dense local ranks (bit 31 = the hash
//                         occurs more than once, i.e. in more than one sketch) in sorted order next to the positions
//   tc_vr_scatter_ranks_dev  received (position, local rank) segments, one per source range -> this rank's slice of the
//                         position-ordered rank array, the source range's base added
//   tc_sketches_upload_ranks_dev  the all-gathered rank array -> resident set (bitmap, union lists, tile flags ...)
namespace {
template <typename A>
__global__ void vr_count_kernel(const A *a, const uint32_t *b, uint64_t m, int mode, const unsigned long long *spl, uint32_t nspl,
                                unsigned long long per, uint32_t nb, unsigned long long *counts)
{
    __shared__ unsigned long long s_c[64];
    if (threadIdx.x < 64) s_c[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t k = 0;
        if (mode == 0) { const unsigned long long x = (unsigned long long)a[i]; for (uint32_t q = 0; q < nspl; q++) k += spl[q] <= x; }
        else k = (uint32_t)(b[i] / per);
        atomicAdd(&s_c[min(k, nb - 1)], 1ULL);
    }
    __syncthreads();
    if (threadIdx.x < nb && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_c[threadIdx.x]);
}
// every block owns a contiguous chunk: bucket histogram in shared memory, ONE global atomic per bucket and block to reserve
// the block's slice of every bucket, then the elements are placed through shared-memory cursors (the order inside a
// bucket is free: the receiver sorts / scatters)
template <typename A>
__global__ void vr_scatter_kernel(const A *a, const uint32_t *b, uint64_t m, int mode, const unsigned long long *spl, uint32_t nspl,
                                  unsigned long long per, uint32_t nb, uint32_t pos_base, unsigned long long *cursor, A *a_out, uint32_t *b_out)
{
    __shared__ unsigned long long s_c[64], s_base[64];
    const uint64_t chunk = (m + gridDim.x - 1) / gridDim.x, lo = (uint64_t)blockIdx.x * chunk, hi = min(m, lo + chunk);
    auto bucket = [&](A x, uint32_t y) -> uint32_t {
        uint32_t k = 0;
        if (mode == 0) { for (uint32_t q = 0; q < nspl; q++) k += spl[q] <= (unsigned long long)x; }
        else k = (uint32_t)(y / per);
        return min(k, nb - 1);
    };
    if (threadIdx.x < 64) s_c[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&s_c[bucket(a[i], b ? b[i] : pos_base + (uint32_t)i)], 1ULL);
    __syncthreads();
    if (threadIdx.x < nb) { s_base[threadIdx.x] = s_c[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], s_c[threadIdx.x]) : 0ULL; s_c[threadIdx.x] = 0; }
    __syncthreads();
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const A x = a[i];
        const uint32_t y = b ? b[i] : pos_base + (uint32_t)i;
        const uint32_t k = bucket(x, y);
        const unsigned long long at = s_base[k] + atomicAdd(&s_c[k], 1ULL);
        a_out[at] = x; b_out[at] = y;
    }
}
__global__ void vr_rank_kernel(const uint32_t *scan, const uint64_t *keys, uint64_t m, uint32_t *rank_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const bool multi = (i > 0 && keys[i - 1] == keys[i]) || (i + 1 < m && keys[i + 1] == keys[i]);
    rank_out[i] = (scan[i] - 1u) | (multi ? 0x80000000u : 0u);
}
__global__ void vr_place_kernel(const uint32_t *pos, const uint32_t *rank, uint64_t m, const unsigned long long *seg_off, const unsigned long long *seg_base,
                                uint32_t nseg, uint32_t pos_lo, uint64_t share_len, uint32_t *out, unsigned int *bad)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t sg = 0;
    while (sg + 1 < nseg && seg_off[sg + 1] <= i) sg++;
    const uint32_t v = rank[i], p = pos[i];
    if (p < pos_lo || (uint64_t)(p - pos_lo) >= share_len) { atomicExch(bad, 1u); return; }
    out[p - pos_lo] = (uint32_t)((v & 0x7FFFFFFFu) + seg_base[sg]) | (v & 0x80000000u);
}
template <typename A>
int vr_partition(tc_ctx *ctx, const A *a_dev, const uint32_t *b_dev, uint64_t m, int mode, const uint64_t *splitters, uint64_t per,
                 uint32_t nb, uint32_t pos_base, A *a_out, uint32_t *b_out, uint64_t *counts)
{
    TC_CHECK(nb >= 1 && nb <= 64, "tc_vr_partition_dev: %u buckets (1..64)", nb);
    TC_CHECK(mode == 0 || (mode == 1 && per > 0 && b_dev), "tc_vr_partition_dev: bad mode / arguments");
    cudaStream_t st = ctx->stream;
    void *ws = nullptr;
    if (tc_scratch(ctx, 3 * 64 * 8, &ws)) return 1;
    unsigned long long *d_spl = static_cast<unsigned long long *>(ws), *d_cnt = d_spl + 64, *d_cur = d_cnt + 64;
    TC_CUDA(cudaMemsetAsync(d_cnt, 0, 64 * 8, st));
    if (mode == 0 && nb > 1) TC_CUDA(cudaMemcpyAsync(d_spl, splitters, (size_t)(nb - 1) * 8, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((m + 255) / 256, (uint64_t)ctx->num_sms * 8));
    vr_count_kernel<A><<<grid, 256, 0, st>>>(a_dev, b_dev, m, mode, d_spl, nb - 1, per, nb, d_cnt);
    unsigned long long h_cnt[64];
    TC_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
    TC_CUDA(cudaStreamSynchronize(st));
    unsigned long long h_cur[64], acc = 0;
    for (uint32_t k = 0; k < nb; k++) { h_cur[k] = acc; acc += h_cnt[k]; counts[k] = h_cnt[k]; }
    TC_CUDA(cudaMemcpyAsync(d_cur, h_cur, (size_t)nb * 8, cudaMemcpyHostToDevice, st));
    vr_scatter_kernel<A><<<grid, 256, 0, st>>>(a_dev, b_dev, m, mode, d_spl, nb - 1, per, nb, pos_base, d_cur, a_out, b_out);
    ctx->launches += 2;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(st));
    return 0;
}
}  // namespace

extern "C" int tc_vr_partition_dev(tc_ctx *ctx, const uint64_t *a_dev, const uint32_t *b_dev, uint64_t m, int mode,
                                   const uint64_t *splitters, uint64_t per, uint32_t nb, uint32_t pos_base,
                                   uint64_t *a_out_dev, uint32_t *b_out_dev, uint64_t *counts)
{
    TC_CHECK(ctx && counts && (m == 0 || (a_dev && a_out_dev && b_out_dev)), "tc_vr_partition_dev: bad arguments");
    TC_CUDA(cudaSetDevice(ctx->device));
    if (m == 0) { for (uint32_t k = 0; k < nb; k++) counts[k] = 0; return 0; }
    return vr_partition<uint64_t>(ctx, a_dev, b_dev, m, mode, splitters, per, nb, pos_base, a_out_dev, b_out_dev, counts);
}
extern "C" int tc_vr_partition32_dev(tc_ctx *ctx, const uint32_t *a_dev, const uint32_t *b_dev, uint64_t m, uint64_t per, uint32_t nb,
                                     uint32_t *a_out_dev, uint32_t *b_out_dev, uint64_t *counts)
{
    TC_CHECK(ctx && counts && (m == 0 || (a_dev && b_dev && a_out_dev && b_out_dev)), "tc_vr_partition32_dev: bad arguments");
    TC_CUDA(cudaSetDevice(ctx->device));
    if (m == 0) { for (uint32_t k = 0; k < nb; k++) counts[k] = 0; return 0; }
    return vr_partition<uint32_t>(ctx, a_dev, b_dev, m, 1, nullptr, per, nb, 0, a_out_dev, b_out_dev, counts);
}

extern "C" int tc_vr_local_ranks_dev(tc_ctx *ctx, const uint64_t *keys_dev, const uint32_t *pos_dev, uint64_t m,
                                     uint32_t *pos_out_dev, uint32_t *rank_out_dev, uint64_t *n_distinct)
{
    TC_CHECK(ctx && n_distinct && (m == 0 || (keys_dev && pos_dev && pos_out_dev && rank_out_dev)), "tc_vr_local_ranks_dev: bad arguments");
    TC_CHECK(m <= 0x7FFFFFFFull, "tc_vr_local_ranks_dev: %llu hashes exceed 2^31 - 1", (unsigned long long)m);
    TC_CUDA(cudaSetDevice(ctx->device));
    *n_distinct = 0;
    if (!m) return 0;
    cudaStream_t st = ctx->stream;
    size_t t1 = 0, t2 = 0;
    TC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                            (uint32_t *)nullptr, (int)m, 0, 64, st));
    TC_CUDA(cub::DeviceScan::InclusiveSum(nullptr, t2, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)m, st));
    const size_t tmp_bytes = (std::max(t1, t2) + 255) & ~(size_t)255;
    const size_t o_scan = (m * 8 + 255) & ~(size_t)255, o_tmp = o_scan + ((m * 4 + 255) & ~(size_t)255);
    void *ws = nullptr;
    if (tc_scratch(ctx, o_tmp + tmp_bytes, &ws)) return 1;
    uint8_t *w8 = static_cast<uint8_t *>(ws);
    uint64_t *skeys = reinterpret_cast<uint64_t *>(w8);
    uint32_t *scan = reinterpret_cast<uint32_t *>(w8 + o_scan);
    size_t tb = tmp_bytes;
    TC_CUDA(cub::DeviceRadixSort::SortPairs(w8 + o_tmp, tb, keys_dev, skeys, pos_dev, pos_out_dev, (int)m, 0, 64, st));
    cub::CountingInputIterator<uint64_t> counting(0);
    cub::TransformInputIterator<uint32_t, HeadFlagOp, cub::CountingInputIterator<uint64_t>> heads(counting, HeadFlagOp{skeys});
    tb = tmp_bytes;
    TC_CUDA(cub::DeviceScan::InclusiveSum(w8 + o_tmp, tb, heads, scan, (int)m, st));
    vr_rank_kernel<<<nblk(m, 256), 256, 0, st>>>(scan, skeys, m, rank_out_dev);
    uint32_t last = 0;
    TC_CUDA(cudaMemcpyAsync(&last, scan + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    ctx->launches += 12;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(st));
    *n_distinct = last;
    return 0;
}

extern "C" int tc_vr_scatter_ranks_dev(tc_ctx *ctx, const uint32_t *pos_dev, const uint32_t *rank_dev, const uint64_t *seg_off,
                                       const uint64_t *seg_base, uint32_t nseg, uint32_t pos_lo, uint64_t share_len, uint32_t *out_dev)
{
    TC_CHECK(ctx && seg_off && seg_base && nseg >= 1 && nseg <= 64, "tc_vr_scatter_ranks_dev: bad arguments");
    TC_CUDA(cudaSetDevice(ctx->device));
    const uint64_t m = seg_off[nseg];
    if (!m) return 0;
    cudaStream_t st = ctx->stream;
    void *ws = nullptr;
    if (tc_scratch(ctx, 3 * 64 * 8, &ws)) return 1;
    unsigned long long *d_off = static_cast<unsigned long long *>(ws), *d_base = d_off + 65;
    unsigned int *d_bad = reinterpret_cast<unsigned int *>(d_base + 65), h_bad = 0;
    TC_CUDA(cudaMemcpyAsync(d_off, seg_off, (size_t)(nseg + 1) * 8, cudaMemcpyHostToDevice, st));
    TC_CUDA(cudaMemcpyAsync(d_base, seg_base, (size_t)nseg * 8, cudaMemcpyHostToDevice, st));
    TC_CUDA(cudaMemsetAsync(d_bad, 0, 4, st));
    vr_place_kernel<<<nblk(m, 256), 256, 0, st>>>(pos_dev, rank_dev, m, d_off, d_base, nseg, pos_lo, share_len, out_dev, d_bad);
    TC_CUDA(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(st));
    TC_CHECK(!h_bad, "tc_vr_scatter_ranks_dev: a position outside [%u, %u + %llu) arrived", pos_lo, pos_lo, (unsigned long long)share_len);
    return 0;
}

extern "C" int tc_sketches_upload_ranks_dev(tc_ctx *ctx, const uint32_t *ranks_flag_dev, const uint64_t *off, uint32_t n)
{
    TC_CHECK(ctx && ranks_flag_dev && off, "tc_sketches_upload_ranks_dev: bad arguments");
    return upload_common(ctx, nullptr, true, off, n, nullptr, ranks_flag_dev);
}

// rows -> d_hashes through two page-locked staging buffers: host threads gather chunk c+1 while chunk c is copied
extern "C" int tc_sketches_upload_rows(tc_ctx *ctx, const uint64_t *const *rows, const uint64_t *sizes, uint32_t n)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(n > 0, "tc_sketches_upload_rows: n == 0");
    reset_sketches(ctx);                 // d_hashes is overwritten below: a failure part-way must not leave the old set's n/ranks
    std::vector<uint64_t> off((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; i++) off[i + 1] = off[i] + sizes[i];
    const uint64_t total = off[n];
    TC_CHECK(total <= 0x7FFFFFFFull, "tc_sketches_upload_rows: %llu hashes exceed the 2^31 - 1 one context ranks at once", (unsigned long long)total);
    if (ensure_cap(&ctx->d_hashes, &ctx->cap_hashes, std::max<uint64_t>(total, 1))) return 1;
    constexpr uint64_t CHUNK = 4u << 20;                         // hashes per staging buffer (32 MB)
    for (int b = 0; b < 2; b++) {
        if (!ctx->h_stage[b]) TC_CUDA(cudaMallocHost(&ctx->h_stage[b], CHUNK * 8));
        if (!ctx->ev_stage[b]) TC_CUDA(cudaEventCreateWithFlags(&ctx->ev_stage[b], cudaEventDisableTiming));
    }
    const unsigned nthreads = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
    uint32_t g = 0;                                              // sketch holding element `pos`
    int buf = 0;
    for (uint64_t pos = 0; pos < total; pos += CHUNK, buf ^= 1) {
        const uint64_t len = std::min(CHUNK, total - pos);
        TC_CUDA(cudaEventSynchronize(ctx->ev_stage[buf]));       // the copy that last used this buffer is done
        while (off[g + 1] <= pos) g++;
        uint32_t g1 = g;
        while (g1 < n && off[g1] < pos + len) g1++;              // sketches [g, g1) intersect the chunk
        uint64_t *dst = ctx->h_stage[buf];
        auto gather = [&](uint32_t a, uint32_t b) {
            for (uint32_t i = a; i < b; i++) {
                const uint64_t lo = std::max(off[i], pos), hi = std::min(off[i + 1], pos + len);
                if (hi > lo) memcpy(dst + (lo - pos), rows[i] + (lo - off[i]), (hi - lo) * 8);
            }
        };
        const uint32_t cnt = g1 - g;
        if (nthreads == 1 || cnt < 64) gather(g, g1);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nthreads; t++)
                th.emplace_back(gather, g + (uint32_t)((uint64_t)cnt * t / nthreads), g + (uint32_t)((uint64_t)cnt * (t + 1) / nthreads));
            for (std::thread &x : th) x.join();
        }
        TC_CUDA(cudaMemcpyAsync(ctx->d_hashes + pos, dst, len * 8, cudaMemcpyHostToDevice, ctx->stream));
        TC_CUDA(cudaEventRecord(ctx->ev_stage[buf], ctx->stream));
    }
    return upload_common(ctx, ctx->d_hashes, true, off.data(), n);
}

extern "C" int tc_pair_common(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, int impl,
                              uint32_t *out)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->n > 0, "tc_pair_common: no sketches uploaded");
    TC_CHECK(i0 <= i1 && i1 <= ctx->n && j0 <= j1 && j1 <= ctx->n, "tc_pair_common: range out of bounds");
    TC_CHECK(impl != 1 || ctx->hashes_resident, "tc_pair_common: impl 1 reads the original hashes, which an upload from sorted runs does not keep");
    const size_t cells = (size_t)(i1 - i0) * (j1 - j0);
    if (!cells) return 0;
    uint32_t *d_out = nullptr;
    TC_CUDA(cudaMalloc(&d_out, cells * 4));
    int rc = impl == 0 ? tc_launch_tiles_matrix(ctx, i0, i1, j0, j1, d_out)
           : impl == 2 ? tc_launch_join_matrix(ctx, i0, i1, j0, j1, d_out)
           : impl == 3 ? tc_launch_blockjoin_matrix(ctx, i0, i1, j0, j1, d_out)
                       : tc_launch_pair_common_ref(ctx, i0, i1, j0, j1, d_out);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, cells * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { tc_set_error("tc_pair_common: %s", cudaGetErrorString(e)); rc = 1; }
    }
    cudaFree(d_out);
    return rc;
}

extern "C" int tc_pair_common_b(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1,
                                uint32_t sketch_size, uint32_t *out)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->n > 0, "tc_pair_common_b: no sketches uploaded");
    TC_CHECK(sketch_size > 0, "tc_pair_common_b: sketch_size == 0");
    TC_CHECK(i0 <= i1 && i1 <= ctx->n && j0 <= j1 && j1 <= ctx->n, "tc_pair_common_b: range out of bounds");
    const size_t cells = (size_t)(i1 - i0) * (j1 - j0);
    if (!cells) return 0;
    uint32_t *d_out = nullptr;
    TC_CUDA(cudaMalloc(&d_out, cells * 4));
    int rc = tc_launch_tiles_matrix(ctx, i0, i1, j0, j1, d_out, sketch_size);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, cells * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { tc_set_error("tc_pair_common_b: %s", cudaGetErrorString(e)); rc = 1; }
    }
    cudaFree(d_out);
    return rc;
}

// ------------------------------------------------------------------------------ dense pass
// Candidate-edge buffer: min(all pairs, 2^28 entries = 3.2 GB) to start with (TCLUST_EDGE_CAP overrides).  Sets whose
// candidate graph is denser -- e.g. KSSD sketches in a 2^24 hash space at N = 500k (BASELINE configs[4]): ~1.4 % of all
// pairs share a hash, 1.8e9 edges -- make the pass overflow; the kernels keep counting, so the exact need is known
// and the pass is repeated once with a buffer of that size (`need` > 0), memory permitting.  No cudaMemGetInfo on the
// normal path: that call alone costs up to tens of milliseconds.
static int ensure_edges(tc_ctx *ctx, uint64_t need = 0)
{
    if (!need && ctx->d_edges && ctx->edge_cap_for_n == ctx->n) return 0;    // sized once per sketch set
    uint64_t cap = (uint64_t)ctx->n * (ctx->n - 1) / 2;
    uint64_t lim = 1ull << 28;
    const char *env = getenv("TCLUST_EDGE_CAP");                 // hard limit: no growth beyond it
    if (env) lim = strtoull(env, nullptr, 10);
    else if (const char *init = getenv("TCLUST_EDGE_CAP_INIT")) lim = strtoull(init, nullptr, 10);   // first size only
    if (need) {
        TC_CHECK(!env, "tc_allpairs_pass: %llu candidate edges exceed TCLUST_EDGE_CAP=%s", (unsigned long long)need, env);
        size_t free_b = 0, total_b = 0;
        TC_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t want = need + need / 64 + 1024;
        const uint64_t room = (free_b + ctx->edge_cap * sizeof(tc_pair)) / sizeof(tc_pair);   // the old buffer is released first
        TC_CHECK(want <= room - room / 8, "tc_allpairs_pass: %llu candidate edges need %.1f GB, more than the free device memory",
                 (unsigned long long)need, want * sizeof(tc_pair) / 1e9);
        lim = want;
    }
    cap = std::max<uint64_t>(std::min(cap, lim), 1);
    ctx->edge_cap_for_n = ctx->n;
    if (ctx->d_edges && ctx->edge_cap >= cap) return 0;          // kept from an earlier upload
    cudaFree(ctx->d_edges);
    ctx->d_edges = nullptr; ctx->edge_cap = 0;
    TC_CUDA(cudaMalloc(&ctx->d_edges, cap * sizeof(tc_pair)));
    ctx->edge_cap = cap;
    TC_CUDA(cudaMemsetAsync(ctx->d_edge_count, 0, 2 * sizeof(unsigned long long), ctx->stream));
    ctx->n_edges = 0;
    return 0;
}

// defer = true: the pass is only queued on the context's stream -- no count read-back, no overflow handling (the caller
// reads the counters later, tc_k4_collect) -- so that what follows (the device Boruvka) can be queued right behind it.
static int allpairs_pass_impl(tc_ctx *ctx, int radio, uint32_t semb_size, uint32_t part, uint32_t nparts,
                              uint64_t *n_pairs, uint64_t *n_edges, uint32_t start_index = 0, bool defer = false)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->n > 0, "tc_allpairs_pass: no sketches uploaded");
    if (nparts == 0) nparts = 1;
    TC_CHECK(part < nparts, "tc_allpairs_pass: part %u >= nparts %u", part, nparts);
    if (ensure_edges(ctx)) return 1;
    // a pass restarts the candidate list
    TC_CUDA(cudaMemsetAsync(ctx->d_edge_count, 0, 2 * sizeof(unsigned long long), ctx->stream));
    // TCLUST_PASS_IMPL=merge selects the first-generation tiled merge kernel (k3_tile.cu)
    const char *impl = getenv("TCLUST_PASS_IMPL");
    // semantic B needs the ordered merge (the stop rule depends on the union order): merge kernel only
    const bool use_merge = (semb_size > 0 || (impl && !strcmp(impl, "merge"))) && start_index == 0;
    TC_CHECK(!(semb_size > 0 && start_index > 0), "tc_allpairs_pass: start_index is not supported for the semantic-B pass");
    TC_CHECK(start_index <= ctx->n, "tc_allpairs_pass: start_index %u > n %u", start_index, ctx->n);
    ctx->edges_semb_size = semb_size;
    ctx->edges_partial = nparts > 1;
    unsigned long long cnt[2] = {0, 0};
    bool on_shadow = false;               // the pass ran on the reordered copy (k3r_reorder.cu)
    for (int attempt = 0;; attempt++) {
            // default: per work unit the block join (K3J) where it pays, else the per-column join (K3H);
        // TCLUST_PASS_IMPL = join | blockjoin forces one of them, merge the first-generation kernel
        const bool force_join = impl && !strcmp(impl, "join"), force_block = impl && !strcmp(impl, "blockjoin");
        on_shadow = !use_merge && !force_join && start_index == 0 && ctx->use_shadow && ctx->shadow;
        if (on_shadow ? tc_shadow_pass(ctx, radio, part, nparts, force_block)
            : use_merge ? tc_launch_tiles_edges(ctx, radio, part, nparts, semb_size)
            : force_join ? tc_launch_join_edges(ctx, radio, part, nparts, start_index)
                         : tc_launch_blockjoin_edges(ctx, radio, part, nparts, start_index, force_block)) return 1;
        if (defer) return 0;
        TC_CUDA(cudaMemcpyAsync(cnt, ctx->d_edge_count, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (cnt[1] == 0 && cnt[0] <= ctx->edge_cap) break;
        // overflow: cnt[0] is the exact number of candidate edges; grow the buffer to it and repeat the pass once
        TC_CHECK(attempt == 0, "tc_allpairs_pass: %llu candidate edges exceed the device edge buffer (%llu)",
                 cnt[0], (unsigned long long)ctx->edge_cap);
        if (ensure_edges(ctx, cnt[0])) return 1;
        TC_CUDA(cudaMemsetAsync(ctx->d_edge_count, 0, 2 * sizeof(unsigned long long), ctx->stream));
    }
    ctx->n_edges = cnt[0];
    if (n_edges) *n_edges = cnt[0];
    if (n_pairs) {
        if (nparts == 1) *n_pairs = (uint64_t)ctx->n * (ctx->n - 1) / 2 - (uint64_t)start_index * (start_index ? start_index - 1 : 0) / 2;
        else if (use_merge) {
            // pairs j<i inside the tiles owned by this part
            const uint64_t G = ctx->ngroups;
            uint64_t p = 0, t = 0;
            for (uint64_t r = 0; r < G; r++)
                for (uint64_t c = 0; c <= r; c++, t++) {
                    if (t % nparts != part) continue;
                    const uint64_t nr = std::min<uint64_t>(TC_GROUP, ctx->n - r * TC_GROUP);
                    const uint64_t nc = std::min<uint64_t>(TC_GROUP, ctx->n - c * TC_GROUP);
                    p += r == c ? nr * (nr - 1) / 2 : nr * nc;
                }
            *n_pairs = p;
        } else {
            // pairs j<i inside the (row block x column segment) units owned by this part; counted once per
            // (set, start_index, part): the loop is 0.5 ms at N = 100k, a sixteenth of an 8-GPU pass
            const uint64_t ckey = ((uint64_t)part << 48) ^ ((uint64_t)nparts << 33) ^ ((uint64_t)start_index << 1) ^ (on_shadow ? 1u : 0u);
            uint64_t p = ctx->pairs_cache_val;
            if (!ctx->pairs_cache_valid || ctx->pairs_cache_key != ckey) {
                p = 0;
                const std::vector<tc_kh_unit> &all = on_shadow ? ctx->shadow->kh_edge_units : ctx->kh_edge_units;
                for (size_t u = part; u < all.size(); u += nparts) {
                    const tc_kh_unit &un = all[u];
                    for (uint64_t i = std::max<uint64_t>((uint64_t)un.rb * 32, start_index); i < std::min<uint64_t>(ctx->n, (uint64_t)un.rb * 32 + 32); i++) {
                        const uint64_t hi = std::min<uint64_t>(un.c1, i);
                        if (hi > un.c0) p += hi - un.c0;
                    }
                }
                ctx->pairs_cache_key = ckey; ctx->pairs_cache_val = p; ctx->pairs_cache_valid = true;
            }
            *n_pairs = p;
        }
    }
    return 0;
}

extern "C" int tc_allpairs_pass(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, uint64_t *n_pairs,
                                uint64_t *n_edges)
{
    return allpairs_pass_impl(ctx, radio, 0, part, nparts, n_pairs, n_edges);
}

extern "C" int tc_allpairs_pass_from(tc_ctx *ctx, int radio, uint32_t start_index, uint32_t part, uint32_t nparts,
                                     uint64_t *n_pairs, uint64_t *n_edges)
{
    return allpairs_pass_impl(ctx, radio, 0, part, nparts, n_pairs, n_edges, start_index);
}

extern "C" int tc_allpairs_pass_b(tc_ctx *ctx, uint32_t sketch_size, uint32_t part, uint32_t nparts,
                                  uint64_t *n_pairs, uint64_t *n_edges)
{
    TC_CHECK(sketch_size > 0, "tc_allpairs_pass_b: sketch_size == 0");
    return allpairs_pass_impl(ctx, 0, sketch_size, part, nparts, n_pairs, n_edges);
}

extern "C" int tc_edges_download(tc_ctx *ctx, tc_pair *out, uint64_t cap, uint64_t *n)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    if (n) *n = ctx->n_edges;
    const uint64_t m = std::min(cap, ctx->n_edges);
    if (m) {
        TC_CUDA(cudaMemcpyAsync(out, ctx->d_edges, m * sizeof(tc_pair), cudaMemcpyDeviceToHost, ctx->stream));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

// -------------------------------------------------------------------------------- Boruvka
// similarity key: ~bits(J); J computed exactly as the reference's host code does
// (MST.cpp:1489-1492: (double)common / (size0+size1-common); :1505-1507: / min(size0,size1)).
// mode: TC_DIST_MASH_A (0) denom = |A|+|B|-common; TC_DIST_AAF (1) / TC_DIST_CONTAIN_B (3) denom = min size;
// TC_DIST_MASH_B (2) denom = min(sketchSize, |A|+|B|-common_B)  (MinHash.cpp:1079-1097)
__device__ __forceinline__ unsigned long long edge_key(const tc_pair &e, const uint32_t *size, int mode, uint32_t semb)
{
    const uint32_t s0 = size[e.i], s1 = size[e.j];
    const uint32_t uni = s0 + s1 - e.common;
    const uint32_t denom = (mode == 1 || mode == 3) ? min(s0, s1) : mode == 2 ? min(semb, uni) : uni;
    const double J = denom == 0 ? 0.0 : __ddiv_rn((double)e.common, (double)denom);
    return ~(unsigned long long)__double_as_longlong(J);
}

__global__ void boruvka_phase1_kernel(const tc_pair *edges, unsigned long long n_edges, const uint32_t *size,
                                      const int32_t *comp, int containment, uint32_t semb,
                                      unsigned long long *best_key)
{
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const tc_pair p = edges[e];
        const int32_t ci = comp[p.i], cj = comp[p.j];
        if (ci == cj) continue;
        const unsigned long long k = edge_key(p, size, containment, semb);
        atomicMin(&best_key[ci], k);
        atomicMin(&best_key[cj], k);
    }
}
__global__ void boruvka_phase2_kernel(const tc_pair *edges, unsigned long long n_edges, const uint32_t *size,
                                      const int32_t *comp, int containment, uint32_t semb,
                                      const unsigned long long *best_key, unsigned long long *best_edge)
{
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const tc_pair p = edges[e];
        const int32_t ci = comp[p.i], cj = comp[p.j];
        if (ci == cj) continue;
        const unsigned long long k = edge_key(p, size, containment, semb);
        const unsigned long long id = ((unsigned long long)p.i << 32) | p.j;
        if (k == best_key[ci]) atomicMin(&best_edge[ci], id);
        if (k == best_key[cj]) atomicMin(&best_edge[cj], id);
    }
}

extern "C" int tc_boruvka_upload_comp(tc_ctx *ctx, const int32_t *comp, uint32_t n)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(n == ctx->n, "tc_boruvka_upload_comp: n mismatch");
    if (ensure_cap(&ctx->d_comp, &ctx->cap_comp, n)) return 1;
    TC_CUDA(cudaMemcpyAsync(ctx->d_comp, comp, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

static int boruvka_grid(tc_ctx *ctx)
{
    uint64_t b = (ctx->n_edges + 255) / 256;
    uint64_t cap = (uint64_t)ctx->num_sms * 8;
    return (int)std::max<uint64_t>(1, std::min(b, cap));
}

extern "C" int tc_boruvka_phase1(tc_ctx *ctx, int containment, uint64_t *best_key_dev)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->d_comp, "tc_boruvka_phase1: component labels not uploaded");
    TC_CHECK(containment >= 0 && containment <= 3, "tc_boruvka_phase1: distance mode %d unknown", containment);
    TC_CHECK((containment == TC_DIST_MASH_B) == (ctx->edges_semb_size > 0),
             "tc_boruvka_phase1: distance mode %d does not match the pass that produced the edges", containment);
    TC_CUDA(cudaMemsetAsync(best_key_dev, 0xFF, (size_t)ctx->n * 8, ctx->stream));
    if (ctx->n_edges) {
        boruvka_phase1_kernel<<<boruvka_grid(ctx), 256, 0, ctx->stream>>>(
            ctx->d_edges, ctx->n_edges, ctx->d_size, ctx->d_comp, containment, ctx->edges_semb_size,
            reinterpret_cast<unsigned long long *>(best_key_dev));
        ctx->launches++;
    }
    TC_CUDA(cudaGetLastError());
    if (!ctx->boruvka_async) TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tc_boruvka_phase2(tc_ctx *ctx, int containment, const uint64_t *best_key_dev, uint64_t *best_edge_dev)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->d_comp, "tc_boruvka_phase2: component labels not uploaded");
    TC_CUDA(cudaMemsetAsync(best_edge_dev, 0xFF, (size_t)ctx->n * 8, ctx->stream));
    if (ctx->n_edges) {
        boruvka_phase2_kernel<<<boruvka_grid(ctx), 256, 0, ctx->stream>>>(
            ctx->d_edges, ctx->n_edges, ctx->d_size, ctx->d_comp, containment, ctx->edges_semb_size,
            reinterpret_cast<const unsigned long long *>(best_key_dev),
            reinterpret_cast<unsigned long long *>(best_edge_dev));
        ctx->launches++;
    }
    TC_CUDA(cudaGetLastError());
    if (!ctx->boruvka_async) TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tc_boruvka_set_async(tc_ctx *ctx, int on)
{
    TC_CHECK(ctx, "tc_boruvka_set_async: null context");
    ctx->boruvka_async = on != 0;
    return 0;
}

// ------------------------------------------------------------------ host distance helpers
extern "C" int tc_radio(double threshold, int kmer_size)
{
    return (int)(2.0 * std::exp(threshold * (kmer_size - 1)) - 1.0);  // calr(threshold, k-1) truncated
}
extern "C" double tc_mash_dist(int common, int size0, int size1, int kmer_size)
{
    const double inv_kmer_size = 1.0 / kmer_size;
    int denom = size0 + size1 - common;
    double jaccard = denom == 0 ? 0.0 : (double)common / denom;
    if (jaccard == 1.0) return 0.0;
    if (jaccard == 0.0) return 1.0;
    double ratio = (2.0 * jaccard) / (1.0 + jaccard);
    return -inv_kmer_size * std::log(ratio);
}
extern "C" double tc_aaf_dist(int common, int size0, int size1, int kmer_size)
{
    const double inv_kmer_size = 1.0 / kmer_size;
    int denom = size0 < size1 ? size0 : size1;
    double containment = denom == 0 ? 0.0 : (double)common / denom;
    if (containment == 1.0) return 0.0;
    if (containment == 0.0) return 1.0;
    return -inv_kmer_size * std::log(containment);
}
extern "C" double tc_dist_from_key(uint64_t key, int containment, int kmer_size)
{
    const uint64_t bits = ~key;
    double J;
    memcpy(&J, &bits, 8);
    const double inv_kmer_size = 1.0 / kmer_size;
    if (containment == TC_DIST_MASH_B || containment == TC_DIST_CONTAIN_B) {
        // MinHash::distance / containDistance (MinHash.cpp:1102-1153): divide by k, clamp at 1
        if (J == 1.0) return 0.0;
        double d = -std::log(2 * J / (1. + J)) / kmer_size;
        if (d > 1) d = 1;
        return d;
    }
    if (J == 1.0) return 0.0;
    if (J == 0.0) return 1.0;
    if (containment) return -inv_kmer_size * std::log(J);
    double ratio = (2.0 * J) / (1.0 + J);
    return -inv_kmer_size * std::log(ratio);
}
extern "C" double tc_jaccard_b(int common_b, int size0, int size1, int sketch_size)
{
    // MinHash::jaccard (MinHash.cpp:1079-1100): denom = union consumed, completed and capped at sketchSize
    long long denom = (long long)size0 + size1 - common_b;
    if (denom > sketch_size) denom = sketch_size;
    return (double)common_b / (double)denom;
}
extern "C" double tc_distance_b(int common_b, int size0, int size1, int sketch_size, int kmer_size)
{
    // MinHash::distance (MinHash.cpp:1125-1153)
    const double jaccard_ = tc_jaccard_b(common_b, size0, size1, sketch_size);
    if (jaccard_ == 1.0) return 0;
    double distance = -std::log(2 * jaccard_ / (1. + jaccard_)) / kmer_size;
    if (distance > 1) distance = 1;
    return distance;
}
extern "C" double tc_contain_distance(int common, int size0, int size1, int kmer_size)
{
    // MinHash::containDistance over containJaccard (MinHash.cpp:908-979, 1102-1123)
    const double jaccard_ = (double)common / (double)std::min(size0, size1);
    if (jaccard_ == 1.0) return 0;
    double distance = -std::log(2 * jaccard_ / (1. + jaccard_)) / kmer_size;
    if (distance > 1) distance = 1;
    return distance;
}

// ------------------------------------------------------------------ threshold-only clustering
__global__ void uf_init_kernel(uint32_t *p, uint32_t n)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) p[v] = v;
}
__global__ void uf_labels_kernel(const uint32_t *p, int32_t *labels, uint32_t n)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    uint32_t x = v;
    while (p[x] != x) x = p[x];
    labels[v] = (int32_t)x;
}

// smallest J in (0, 1] whose reference distance is <= threshold (distance is non-increasing in J): bisection over the
// bit patterns of the positive doubles with the very function that produces the MST weights; 2.0 if none
static double jmin_for_threshold(double threshold, int mode, int kmer_size)
{
    auto ok = [&](double J) { uint64_t b; memcpy(&b, &J, 8); return tc_dist_from_key(~b, mode, kmer_size) <= threshold; };
    if (!ok(1.0)) return 2.0;
    const double one = 1.0;
    uint64_t lo = 1, hi;
    memcpy(&hi, &one, 8);
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        double v;
        memcpy(&v, &mid, 8);
        if (ok(v)) hi = mid; else lo = mid + 1;
    }
    double v;
    memcpy(&v, &lo, 8);
    return v;
}

// Clusters at `threshold` without building (or storing) the MST: the connected components of the candidate pairs with
// distance <= threshold -- exactly the partition generateForest + generateClusterWithBfs give on the MST
// (src/MST.cpp:77-142: an MST keeps, for every threshold, the connectivity of the graph restricted to edges <= it).
// The dense pass runs with a union-find epilogue: no candidate edge is written, memory is O(n).
// labels[v] = smallest vertex id of v's cluster over the pairs of `part`; merge the parts with tc_labels_merge.
extern "C" int tc_cluster_threshold(tc_ctx *ctx, int kmer_size, int containment, double threshold, uint32_t part, uint32_t nparts,
                                    int32_t *labels, uint32_t *n_clusters)
{
    TC_CHECK(ctx && labels, "tc_cluster_threshold: bad arguments");
    TC_CHECK(containment == 0 || containment == 1, "tc_cluster_threshold: containment must be 0 or 1");
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->n > 0, "tc_cluster_threshold: no sketches uploaded");
    const char *impl = getenv("TCLUST_PASS_IMPL");
    TC_CHECK(!(impl && !strcmp(impl, "merge")), "tc_cluster_threshold: not available on the first-generation merge kernel");
    const uint32_t n = ctx->n;
    if (ensure_cap(&ctx->d_uf, &ctx->cap_uf, (uint64_t)2 * n)) return 1;
    cudaStream_t st = ctx->stream;
    uf_init_kernel<<<nblk(n, 256), 256, 0, st>>>(ctx->d_uf, n);
    ctx->launches++;
    ctx->uf_parent = ctx->d_uf;
    ctx->uf_jmin = jmin_for_threshold(threshold, containment, kmer_size);
    ctx->uf_containment = containment;
    ctx->uf_perm = nullptr;
    if (ctx->shadow) {
        ctx->shadow->uf_parent = ctx->d_uf; ctx->shadow->uf_jmin = ctx->uf_jmin; ctx->shadow->uf_containment = containment;
        ctx->shadow->uf_perm = ctx->shadow->d_perm;
    }
    uint64_t np = 0, ne = 0;
    const int rc = allpairs_pass_impl(ctx, tc_radio(threshold, kmer_size), 0, part, nparts, &np, &ne);
    ctx->uf_parent = nullptr;
    if (ctx->shadow) ctx->shadow->uf_parent = nullptr;
    ctx->edges_partial = true; ctx->n_edges = 0;
    if (rc) return rc;
    int32_t *d_lab = reinterpret_cast<int32_t *>(ctx->d_uf + n);
    uf_labels_kernel<<<nblk(n, 256), 256, 0, st>>>(ctx->d_uf, d_lab, n);
    ctx->launches++;
    TC_CUDA(cudaMemcpyAsync(labels, d_lab, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    TC_CUDA(cudaStreamSynchronize(st));
    if (n_clusters) {
        uint32_t c = 0;
        for (uint32_t v = 0; v < n; v++) c += labels[v] == (int32_t)v;
        *n_clusters = c;
    }
    return 0;
}

// host: into[] := labels of the union of the two partitions (both label arrays: smallest vertex id of the cluster)
extern "C" int tc_labels_merge(int32_t *into, const int32_t *other, uint32_t n, uint32_t *n_clusters)
{
    TC_CHECK(into && other, "tc_labels_merge: bad arguments");
    std::vector<int32_t> p(n);
    std::iota(p.begin(), p.end(), 0);
    auto find = [&](int32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; };
    auto unite = [&](int32_t a, int32_t b) { a = find(a); b = find(b); if (a != b) { if (a < b) p[b] = a; else p[a] = b; } };
    for (uint32_t v = 0; v < n; v++) {
        TC_CHECK(into[v] >= 0 && (uint32_t)into[v] < n && other[v] >= 0 && (uint32_t)other[v] < n, "tc_labels_merge: label out of range at %u", v);
        unite((int32_t)v, into[v]); unite((int32_t)v, other[v]);
    }
    uint32_t c = 0;
    for (uint32_t v = 0; v < n; v++) { into[v] = find((int32_t)v); c += into[v] == (int32_t)v; }
    if (n_clusters) *n_clusters = c;
    return 0;
}

// ------------------------------------------------------------------------------ MST driver
namespace {
struct HostUF {
    std::vector<int32_t> p;
    explicit HostUF(uint32_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int32_t find(int32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
};
struct NoSyncPass {                                      // the launchers of the dense pass only queue while this lives
    tc_ctx *c; bool was;
    explicit NoSyncPass(tc_ctx *x) : c(x), was(x->pass_nosync) { x->pass_nosync = true; if (x->shadow) x->shadow->pass_nosync = true; }
    ~NoSyncPass() { c->pass_nosync = was; if (c->shadow) c->shadow->pass_nosync = was; }
};
}  // namespace

static bool edges_can_grow(tc_ctx *ctx, uint64_t need)
{
    if (getenv("TCLUST_EDGE_CAP")) return false;                 // hard limit on the edge buffer
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return false; }
    const uint64_t want = need + need / 64 + 1024;
    const uint64_t room = (free_b + ctx->edge_cap * sizeof(tc_pair)) / sizeof(tc_pair);
    return want <= room - room / 8;
}

// Bounded edge memory: the candidate graph does not fit the edge buffer (a set in which
// most sketches are related has ~N^2/2 candidate pairs: 60 GB at N = 100k).  The reference never materialises it either:
// every task of 8 rows reduces its edges to a forest and the forests are merged at the end (src/MST.cpp:1373-1386,
// 571-572, 1715-1723).  Same identity here, batch-wise: MSF(E1 u E2) = MSF(MSF(E1) u E2).  The work units of the dense
// pass are launched in batches whose worst-case edge count fits the buffer; whenever the next batch might not fit, the
// device Boruvka reduces (forest so far + buffered edges) to a forest of <= n-1 edges that stays on the device and the
// buffer restarts empty.  Exact: the strict order (key, i, j) makes the minimum spanning forest unique.
static int mst_bounded(tc_ctx *ctx, int kmer_size, int mode, int radio, uint32_t start_index, std::vector<tc_edge> &mst,
                       std::vector<int32_t> *comp)
{
    const char *impl = getenv("TCLUST_PASS_IMPL");
    TC_CHECK(!(impl && (!strcmp(impl, "merge") || !strcmp(impl, "join"))),
             "tc_mst: the candidate edges exceed the edge buffer and TCLUST_PASS_IMPL=%s has no bounded-memory form", impl);
    const bool force_block = impl && !strcmp(impl, "blockjoin");
    const bool on_shadow = start_index == 0 && ctx->use_shadow && ctx->shadow;
    tc_ctx *src = on_shadow ? ctx->shadow : ctx;
    if (on_shadow) { src->d_edges = ctx->d_edges; src->edge_cap = ctx->edge_cap; src->d_kh_state = ctx->d_kh_state; }
    if (tc_launch_blockjoin_edges(src, radio, 0, 1, start_index, force_block, true)) return 1;
    cudaStream_t st = ctx->stream;
    TC_CUDA(cudaMemsetAsync(ctx->d_edge_count, 0, 2 * sizeof(unsigned long long), st));
    if (tc_k4_reserve(ctx, ctx->n)) return 1;
    uint64_t count = 0, batches = 0, folds = 0;
    uint32_t nfin = 0;
    auto fin_key = [&]() -> const unsigned long long * { return nfin ? ctx->k4.d_fin : nullptr; };
    auto fin_id = [&]() -> const unsigned long long * { return nfin ? ctx->k4.d_fin + ctx->n : nullptr; };
    for (int blk = 0; blk < 3; blk++) {              // sparse tiles, block-join units, hash-join units
        const size_t nu = tc_blockjoin_list_size(src, blk);
        for (size_t u = 0; u < nu;) {
            size_t u1 = u;
            uint64_t bound = 0;
            while (u1 < nu && count + bound + tc_blockjoin_unit_bound(src, blk, u1) <= ctx->edge_cap) bound += tc_blockjoin_unit_bound(src, blk, u1++);
            if (u1 == u) {
                TC_CHECK(count > 0, "tc_mst: one work unit of the dense pass can emit %llu candidate edges, more than the edge buffer holds (%llu): raise TCLUST_EDGE_CAP",
                         (unsigned long long)tc_blockjoin_unit_bound(src, blk, u), (unsigned long long)ctx->edge_cap);
                if (tc_k4_enqueue(ctx, mode, fin_key(), fin_id(), nfin) || tc_k4_fold(ctx, &nfin)) return 1;
                count = 0; folds++;
                continue;
            }
            if (tc_blockjoin_run_range(src, radio, start_index, blk, u, u1)) return 1;
            if (on_shadow) {
                ctx->d_kh_state = src->d_kh_state;
                if (tc_shadow_remap_from(ctx, count)) return 1;
            }
            unsigned long long cnt[2] = {0, 0};
            TC_CUDA(cudaMemcpyAsync(cnt, ctx->d_edge_count, sizeof cnt, cudaMemcpyDeviceToHost, st));
            TC_CUDA(cudaStreamSynchronize(st));
            TC_CHECK(cnt[1] == 0 && cnt[0] <= ctx->edge_cap, "tc_mst: a batch of the bounded-memory pass overflowed the edge buffer (%llu > %llu)",
                     cnt[0], (unsigned long long)ctx->edge_cap);
            count = cnt[0];
            u = u1; batches++;
        }
    }
    ctx->launches += src != ctx ? src->launches : 0;
    if (src != ctx) src->launches = 0;
    if (tc_k4_enqueue(ctx, mode, fin_key(), fin_id(), nfin)) return 1;
    uint64_t cnt = 0;
    bool overflow = false;
    if (tc_k4_collect(ctx, mode, kmer_size, mst, &cnt, &overflow, comp)) return 1;
    TC_CHECK(!overflow, "tc_mst: the last batch of the bounded-memory pass overflowed the edge buffer");
    ctx->n_edges = count; ctx->edges_partial = true;
    ctx->bounded_batches = batches; ctx->bounded_folds = folds;
    if (getenv("TCLUST_TIMING")) fprintf(stderr, "[tclust timing] bounded-memory MST: %llu batches, %llu forest reductions, edge buffer %llu entries\n",
                                         (unsigned long long)batches, (unsigned long long)folds + 1, (unsigned long long)ctx->edge_cap);
    return 0;
}

// mode = distance mode of the Boruvka keys; complete = add weight-1 edges between the components that share no
// hash (the brute-force reference builds the MST of the complete graph).
// The dense pass (this GPU's part) and the whole Boruvka (k4_boruvka.cu; with a multi-GPU group the per-round
// reduction over peer memory) are queued back to back; the host waits once, for the sorted forest.
static int mst_impl(tc_ctx *ctx, int kmer_size, int mode, int radio, uint32_t semb_size, bool complete,
                    tc_edge *out, uint64_t *n_out, uint32_t start_index = 0, const tc_edge *pre_mst = nullptr,
                    uint64_t n_pre = 0, bool edges_only = false)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->n > 0, "tc_mst: no sketches uploaded");
    const uint32_t n = ctx->n;
    const uint32_t part = ctx->k4.nranks > 1 ? (uint32_t)ctx->k4.rank : 0u, nparts = ctx->k4.nranks > 1 ? (uint32_t)ctx->k4.nranks : 1u;
    std::vector<tc_edge> mst;
    std::vector<int32_t> comp;
    ctx->bounded_batches = ctx->bounded_folds = 0;
    // the plain call (no completion with distance-1 edges, no old MST to merge): the forest is converted straight into the
    // caller's array (room for n - 1 edges by contract), without a vector in between
    tc_edge *direct = (out && !complete && !n_pre) ? out : nullptr;
    uint64_t n_direct = 0;
    bool is_direct = direct != nullptr;
    for (int attempt = 0;; attempt++) {
        uint64_t cnt = 0;
        bool overflow = false;
        if (!edges_only) {
            NoSyncPass nosync(ctx);
            if (allpairs_pass_impl(ctx, radio, semb_size, part, nparts, nullptr, nullptr, start_index, true)) return 1;
        }
        if (tc_k4_enqueue(ctx, mode, nullptr, nullptr, 0)) return 1;
        if (tc_k4_collect(ctx, mode, kmer_size, mst, &cnt, &overflow, complete ? &comp : nullptr, direct, &n_direct)) return 1;
        ctx->n_edges = std::min<uint64_t>(cnt, ctx->edge_cap);
        if (!overflow) break;
        TC_CHECK(!edges_only, "tc_mst_of_edges: the candidate edges of the last pass overflowed the edge buffer");
        // the candidate edges did not fit: cnt is the exact need; grow the buffer to it and repeat once.  In a multi-GPU
        // group the kernel has told every rank, so all of them repeat the (collective) call together.
        TC_CHECK(attempt == 0, "tc_mst: %llu candidate edges exceed the device edge buffer (%llu)",
                 (unsigned long long)cnt, (unsigned long long)ctx->edge_cap);
        if (cnt > ctx->edge_cap && !edges_can_grow(ctx, cnt)) {
            // no room for the whole candidate graph: build the MST in bounded edge memory instead
            TC_CHECK(nparts == 1, "tc_mst: rank %u of the multi-GPU group cannot hold its %llu candidate edges (buffer %llu)", part,
                     (unsigned long long)cnt, (unsigned long long)ctx->edge_cap);
            TC_CHECK(semb_size == 0 && !edges_only, "tc_mst: %llu candidate edges exceed the device edge buffer (%llu)",
                     (unsigned long long)cnt, (unsigned long long)ctx->edge_cap);
            if (mst_bounded(ctx, kmer_size, mode, radio, start_index, mst, complete ? &comp : nullptr)) return 1;
            is_direct = false;
            break;
        }
        if (cnt > ctx->edge_cap && ensure_edges(ctx, cnt)) return 1;      // (in a group: only the ranks that overflowed grow)
    }
    auto by_weight = [](const tc_edge &x, const tc_edge &y) {
        if (x.dist != y.dist) return x.dist < y.dist;
        if (x.pre != y.pre) return x.pre < y.pre;
        return x.suf < y.suf;
    };
    if (complete) {
        // pairs without a shared hash have distance exactly 1 (J = 0): chain the remaining components
        std::vector<char> seen(n, 0);
        const int32_t first_root = comp[0];
        seen[first_root] = 1;
        bool added = false;
        for (uint32_t v = 1; v < n; v++) {
            const int32_t r = comp[v];
            if (seen[r]) continue;
            seen[r] = 1;
            mst.push_back(tc_edge{(int32_t)v, 0, 1.0});
            added = true;
        }
        if (added) std::sort(mst.begin(), mst.end(), by_weight);
    }
    if (n_pre) {
        // append: MST(old graph + new edges) = MST(old MST + forest of the new edges)  (sub_command.cpp:1700-1706)
        std::vector<tc_edge> all(pre_mst, pre_mst + n_pre);
        for (const tc_edge &e : all)
            TC_CHECK(e.pre >= 0 && e.suf >= 0 && (uint32_t)e.pre < n && (uint32_t)e.suf < n, "tc_mst_append: pre-MST edge out of range");
        all.insert(all.end(), mst.begin(), mst.end());
        std::sort(all.begin(), all.end(), by_weight);
        HostUF uf2(n);
        mst.clear();
        for (const tc_edge &e : all) {
            const int32_t a = uf2.find(e.pre), b = uf2.find(e.suf);
            if (a == b) continue;
            uf2.p[a] = b;
            mst.push_back(e);
        }
    }
    if (is_direct) { if (n_out) *n_out = n_direct; return 0; }
    if (out) std::copy(mst.begin(), mst.end(), out);
    if (n_out) *n_out = mst.size();
    return 0;
}

extern "C" int tc_mst(tc_ctx *ctx, int kmer_size, int containment, double threshold, tc_edge *out, uint64_t *n_out)
{
    TC_CHECK(containment == 0 || containment == 1, "tc_mst: containment must be 0 or 1");
    return mst_impl(ctx, kmer_size, containment, tc_radio(threshold, kmer_size), 0, false, out, n_out);
}

// Boruvka alone over the candidate edges the last tc_allpairs_pass* left on the device (in a group: every rank's own
// part, reduced inside the kernel)
extern "C" int tc_mst_of_edges(tc_ctx *ctx, int kmer_size, int containment, tc_edge *out, uint64_t *n_out)
{
    TC_CHECK(containment >= 0 && containment <= 3, "tc_mst_of_edges: distance mode %d unknown", containment);
    TC_CHECK((containment == TC_DIST_MASH_B) == (ctx->edges_semb_size > 0),
             "tc_mst_of_edges: distance mode %d does not match the pass that produced the edges", containment);
    return mst_impl(ctx, kmer_size, containment, 0, ctx->edges_semb_size, false, out, n_out, 0, nullptr, 0, true);
}

extern "C" int tc_mst_append(tc_ctx *ctx, int kmer_size, int containment, double threshold, uint32_t start_index,
                             const tc_edge *pre_mst, uint64_t n_pre, tc_edge *out, uint64_t *n_out)
{
    TC_CHECK(containment == 0 || containment == 1, "tc_mst_append: containment must be 0 or 1");
    return mst_impl(ctx, kmer_size, containment, tc_radio(threshold, kmer_size), 0, false, out, n_out, start_index, pre_mst, n_pre);
}

extern "C" int tc_mst_bruteforce(tc_ctx *ctx, int kmer_size, uint32_t sketch_size, int containment, tc_edge *out,
                                 uint64_t *n_out)
{
    if (containment)   // containDistance: full intersection / min size, no size-ratio filter
        return mst_impl(ctx, kmer_size, TC_DIST_CONTAIN_B, 0x7fffffff, 0, true, out, n_out);
    TC_CHECK(sketch_size > 0, "tc_mst_bruteforce: sketch_size == 0");
    return mst_impl(ctx, kmer_size, TC_DIST_MASH_B, 0, sketch_size, true, out, n_out);
}
