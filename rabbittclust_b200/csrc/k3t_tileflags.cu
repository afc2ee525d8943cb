// k3t_tileflags.cu -- which 32 x 32 tiles of the all-pairs space can hold a candidate pair at all
//
// The reference never looks at a pair that shares no hash: compute_minhash_mst walks, for every hash of genome i, the
// posting list of that hash in its inverted index (src/MST.cpp:1409-1435, index src/SketchInfo.h:95-161).  The dense
// pass (k3h_join.cu) evaluates 32 x 32 tiles; this file gives it the same selectivity at tile granularity:
//
//   * k3h_distinct_kernel has left, per block of 32 sketches, the ascending list of the block's distinct ranks that
//     occur in at least two sketches of the set (the union lists).  Concatenated and sorted by rank they ARE the
//     inverted index at block granularity: rank -> blocks that hold it (k3t_gather_kernel + one radix sort of
//     (rank, block) pairs -- 4 % of the elements in the bench set).
//   * k3t_flag_kernel walks every posting list and sets bit (row block, column block) for every pair of blocks on it:
//     a tile whose bit stays clear has no rank in common between its 32 rows and its 32 columns, i.e. all 1024
//     intersections are empty and the tile produces no candidate edge (common >= 1 is the reference's candidate rule,
//     src/MST.cpp:1470).  Diagonal tiles are always flagged; blocks without a union list flag their row and column.
//   * k3t_rowcount_kernel counts the flagged tiles of every row block (unit ordering / routing on the host).
//
// Exactness: the flags only remove tiles in which every pair has common == 0; `common` of every remaining pair is
// computed by the same kernels as before.
#include "tc_common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

constexpr uint32_t KT_DREC = 8;          // record stride of d_bu_count (KH_DREC in k3h_join.cu)

// one CTA per block: copy the block's union list behind the lists of the earlier blocks, tagging every key with the block
__global__ void k3t_gather_kernel(const uint32_t *bu_keys, const uint32_t *bu_count, const unsigned long long *off,
                                  const unsigned long long *pref, uint32_t nrb, uint32_t *okey, uint32_t *oblk)
{
    const uint32_t b = blockIdx.x;
    if (b >= nrb) return;
    const uint32_t D = bu_count[b * KT_DREC];
    if (D == 0xFFFFFFFFu) return;
    const uint32_t *src = bu_keys + off[b * 32];
    const unsigned long long dst = pref[b];
    for (uint32_t d = threadIdx.x; d < D; d += blockDim.x) { okey[dst + d] = src[d]; oblk[dst + d] = b; }
}

// postings sorted by (rank, block): entry i pairs up with every earlier entry of the same rank (an earlier block)
// (*work > limit: the guard below has decided to flag every tile instead, nothing to do here)
__global__ void k3t_flag_kernel(const uint32_t *key, const uint32_t *blk, unsigned long long m, uint32_t *flags, uint32_t words,
                                const unsigned long long *work, unsigned long long limit)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || i == 0 || *work > limit) return;
    const uint32_t k = key[i], b = blk[i];
    uint32_t *row = flags + (size_t)b * words;
    for (unsigned long long j = i; j-- > 0;) {
        if (key[j] != k) break;
        const uint32_t c = blk[j], bit = 1u << (c & 31);
        if (!(row[c >> 5] & bit)) atomicOr(&row[c >> 5], bit);
    }
}

// work of k3t_flag_kernel = sum over the posting lists of L(L-1)/2: the last entry of every list walks back to its start
// (O(postings) in total) and adds its list's share
__global__ void k3t_work_kernel(const uint32_t *key, unsigned long long m, unsigned long long *work)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (i < m && !(i + 1 < m && key[i + 1] == key[i])) {
        unsigned long long L = 1;
        for (unsigned long long j = i; j-- > 0 && key[j] == key[i];) L++;
        v = L * (L - 1) / 2;
    }
    // one atomic per warp: a million single-address atomics cost more than the rest of the flag build
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(work, v);
}

// the guard's other branch: every tile flagged (a superset is always correct)
__global__ void k3t_saturate_kernel(uint32_t *flags, unsigned long long fwords, const unsigned long long *work, unsigned long long limit)
{
    if (*work <= limit) return;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < fwords;
         i += (unsigned long long)gridDim.x * blockDim.x) flags[i] = 0xFFFFFFFFu;
}

// diagonal tiles, and the row + column of every block without a union list (K3H streams those without a list)
__global__ void k3t_fixed_kernel(const uint32_t *bu_count, uint32_t nrb, uint32_t *flags, uint32_t words)
{
    const uint32_t b = blockIdx.x;
    if (b >= nrb) return;
    if (threadIdx.x == 0) atomicOr(&flags[(size_t)b * words + (b >> 5)], 1u << (b & 31));
    if (bu_count[b * KT_DREC] != 0xFFFFFFFFu) return;
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) flags[(size_t)b * words + w] = 0xFFFFFFFFu;
    for (uint32_t r = b + 1 + threadIdx.x; r < nrb; r += blockDim.x) atomicOr(&flags[(size_t)r * words + (b >> 5)], 1u << (b & 31));
}

__global__ void k3t_rowcount_kernel(const uint32_t *flags, uint32_t nrb, uint32_t words, uint32_t *count)
{
    const uint32_t rb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (rb >= nrb) return;
    uint32_t c = 0;
    for (uint32_t w = lane; w <= (rb >> 5); w += 32) {
        uint32_t v = flags[(size_t)rb * words + w];
        if (w == (rb >> 5) && (rb & 31) != 31) v &= (2u << (rb & 31)) - 1u;     // columns <= rb only
        c += __popc(v);
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) count[rb] = c;
}

}  // namespace

// dist = host copy of d_bu_count (tc_kh_prepare has just downloaded it).  Leaves ctx->d_tile_flags / tile_words /
// tile_rowcount (host) valid, or tile_words == 0 when the flags are switched off (TCLUST_TILE_FLAGS=0).
int tc_tile_flags_build(tc_ctx *ctx, const std::vector<uint32_t> &dist)
{
    ctx->tile_words = 0;
    ctx->tile_rowcount.clear();
    ctx->tiles_flagged = 0;
    if (const char *e = getenv("TCLUST_TILE_FLAGS")) if (!strcmp(e, "0")) return 0;
    const uint32_t nrb = (ctx->n + 31) / 32;
    if (nrb == 0) return 0;
    cudaStream_t st = ctx->stream;
    const uint32_t words = nrb / 32 + 2;                      // one spare word: the kernels fetch 8 bits with a funnel shift
    const uint64_t fwords = (uint64_t)nrb * words;
    if (fwords > ctx->cap_tile_flags) {
        cudaFree(ctx->d_tile_flags); ctx->d_tile_flags = nullptr; ctx->cap_tile_flags = 0;
        TC_CUDA(cudaMalloc(&ctx->d_tile_flags, fwords * 4));
        ctx->cap_tile_flags = fwords;
    }
    if (nrb > ctx->cap_tile_rowcount) {
        cudaFree(ctx->d_tile_rowcount); ctx->d_tile_rowcount = nullptr; ctx->cap_tile_rowcount = 0;
        TC_CUDA(cudaMalloc(&ctx->d_tile_rowcount, (size_t)nrb * 4));
        ctx->cap_tile_rowcount = nrb;
    }
    TC_CUDA(cudaMemsetAsync(ctx->d_tile_flags, 0, fwords * 4, st));
    // ---- postings: the union lists back to back
    std::vector<unsigned long long> pref((size_t)nrb + 1, 0);
    for (uint32_t b = 0; b < nrb; b++) {
        const uint32_t D = dist[(size_t)b * KT_DREC];
        pref[b + 1] = pref[b] + (D == 0xFFFFFFFFu ? 0u : D);
    }
    const unsigned long long m = pref[nrb];
    TC_CHECK(m <= 0x7FFFFFFFull, "tc_tile_flags_build: %llu postings", m);
    if (m > 1) {
        int end_bit = 1;
        while (end_bit < 32 && (ctx->total >> end_bit)) end_bit++;           // ranks < total
        size_t tmp_bytes = 0;
        TC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                                (uint32_t *)nullptr, (int)m, 0, end_bit, st));
        const size_t a256 = 255;
        const size_t o_pref = 0, o_k0 = (o_pref + (nrb + 2) * 8 + a256) & ~a256, o_b0 = (o_k0 + m * 4 + a256) & ~a256;
        const size_t o_k1 = (o_b0 + m * 4 + a256) & ~a256, o_b1 = (o_k1 + m * 4 + a256) & ~a256, o_tmp = (o_b1 + m * 4 + a256) & ~a256;
        void *ws = nullptr;
        if (tc_scratch(ctx, o_tmp + tmp_bytes, &ws)) return 1;
        uint8_t *w8 = static_cast<uint8_t *>(ws);
        unsigned long long *d_pref = reinterpret_cast<unsigned long long *>(w8 + o_pref);
        uint32_t *k0 = reinterpret_cast<uint32_t *>(w8 + o_k0), *b0 = reinterpret_cast<uint32_t *>(w8 + o_b0);
        uint32_t *k1 = reinterpret_cast<uint32_t *>(w8 + o_k1), *b1 = reinterpret_cast<uint32_t *>(w8 + o_b1);
        TC_CUDA(cudaMemcpyAsync(d_pref, pref.data(), (size_t)(nrb + 1) * 8, cudaMemcpyHostToDevice, st));
        k3t_gather_kernel<<<nrb, 256, 0, st>>>(ctx->d_bu_keys, ctx->d_bu_count, reinterpret_cast<const unsigned long long *>(ctx->d_off),
                                              d_pref, nrb, k0, b0);
        TC_CUDA(cub::DeviceRadixSort::SortPairs(w8 + o_tmp, tmp_bytes, (const uint32_t *)k0, k1, (const uint32_t *)b0, b1, (int)m, 0, end_bit, st));
        // The flag kernel visits every pair of blocks of every posting list.  A rank held by most blocks (one species: every
        // shared rank is in every block) makes that quadratic in the number of blocks; when the visits exceed what a pass over
        // ALL tiles would cost anyway, every tile is flagged instead (a superset is always correct).
        unsigned long long *d_work = reinterpret_cast<unsigned long long *>(w8 + o_pref) + (nrb + 1), h_work = 0;
        TC_CUDA(cudaMemsetAsync(d_work, 0, 8, st));
        k3t_work_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(k1, m, d_work);
        // the decision is taken on the device (both kernels read the estimate): no host round trip in the upload's tail
        const unsigned long long tiles = (unsigned long long)nrb * (nrb + 1) / 2;
        const unsigned long long limit = std::max<unsigned long long>(64ull * tiles, 1ull << 28);
        k3t_flag_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(k1, b1, m, ctx->d_tile_flags, words, d_work, limit);
        k3t_saturate_kernel<<<(unsigned)std::min<unsigned long long>((fwords + 255) / 256, 4096), 256, 0, st>>>(ctx->d_tile_flags, fwords, d_work, limit);
        if (getenv("TCLUST_TIMING")) {
            TC_CUDA(cudaMemcpyAsync(&h_work, d_work, 8, cudaMemcpyDeviceToHost, st));
            TC_CUDA(cudaStreamSynchronize(st));
            if (h_work > limit) fprintf(stderr, "[tclust timing] tile flags: %llu block-pair visits for %llu tiles: every tile flagged\n", h_work, tiles);
        }
        ctx->launches += 3 + (end_bit + 7) / 8;
    }
    k3t_fixed_kernel<<<nrb, 128, 0, st>>>(ctx->d_bu_count, nrb, ctx->d_tile_flags, words);
    k3t_rowcount_kernel<<<(nrb * 32 + 255) / 256, 256, 0, st>>>(ctx->d_tile_flags, nrb, words, ctx->d_tile_rowcount);
    ctx->launches += 2;
    ctx->tile_rowcount.resize(nrb);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->tile_rowcount.data(), ctx->d_tile_rowcount, (size_t)nrb * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // `pref` (host) must outlive its copy as well
    if (e != cudaSuccess) { tc_set_error("tc_tile_flags_build: %s", cudaGetErrorString(e)); return 1; }
    for (uint32_t c : ctx->tile_rowcount) ctx->tiles_flagged += c;
    ctx->tile_words = words;
    return 0;
}
