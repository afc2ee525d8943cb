// k1_sketch.cu -- K1: MinHash sketch construction (HP-1).
//
// Replaces Sketch::MinHash::update + MinHashHeap::tryInsert + heapToList of the reference
// (RabbitSketch/src/MinHash.cpp:465-747, 287-348, 756-843; Murmur3 RabbitSketch/src/
// MurmurHash3.cpp:259-336; reverseComplement MinHash.cpp:1185-1195).  Semantics reproduced
// bit-exactly for arbitrary input bytes:
//   * bytes 'a'..'z' are upper-cased (MinHash.cpp:472-479);
//   * reverse complement of byte c is "TGAC"[(c>>1)&3] (so N -> C etc.);
//   * canonical k-mer = memcmp-smaller of forward and reverse-complement strand (AVX2 build,
//     MinHash.cpp:667-670,696), forward on ties;
//   * hash = MurmurHash3_x64_128(k ASCII bytes of that strand, seed)[0];
//   * sketch = the s smallest DISTINCT hashes, ascending.
//
// Design (DESIGN.md "K1"): the path is integer-ALU bound (~125 INT32 ops per k-mer), not
// HBM bound.  A genome is cut into segments of whole windows; one CTA sketches one segment:
// every thread walks its own contiguous run of windows with the two strands kept as
// big-endian ASCII words in registers (shift one byte per base), so the memcmp is a
// multi-word integer compare and there is no special path for non-ACGT bytes.  Hashes not
// above the running threshold tau (the s-th smallest distinct hash seen so far) are appended
// to a shared-memory candidate buffer which is sorted/deduplicated/truncated (bitonic) when
// it fills; tau is shared between the segments of a genome through a global atomicMin.  A
// second kernel merges the per-segment partial sketches of each genome.
#include "tc_common.cuh"
#include "k_sort.cuh"
#include <algorithm>

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_STEP = 8;                       // bases per thread and load (one aligned 8-byte word)
constexpr int K1_VOTE = 4;                       // bases per thread between block-wide buffer checks
constexpr int K1_STEP_CAND = K1_THREADS * K1_VOTE;   // most candidates appended between two checks

struct K1Seg {
    unsigned long long seq_start;   // byte offset of the first base of the first window
    uint32_t nwin;                  // windows in this segment
    uint32_t genome;
    unsigned long long out_off;     // element offset of this segment's partial sketch
};

struct K1Args {
    const uint8_t *seq;
    const K1Seg *segs;
    uint32_t nseg;
    int k;
    uint32_t seed;
    const uint32_t *sketch_size;     // per genome
    unsigned long long *tau;         // per genome, init UINT64_MAX
    uint64_t *partial;               // per-segment partial sketches
    uint32_t *partial_n;
    unsigned int *seg_counter;
    uint32_t cap;                    // candidate buffer capacity (power of two)
    uint64_t *gscratch;              // per-CTA global candidate buffers when cap does not fit shared memory
    int use_global;
};

// ------------------------------------------------------------------------------- Murmur3
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
__device__ __forceinline__ uint64_t bswap64(uint64_t x)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}

// MurmurHash3_x64_128 (h1 only) of k bytes given as NW little-endian words, zero padded.
template <int NW>
__device__ __forceinline__ uint64_t murmur_h1(const uint64_t (&w)[NW], int k, uint32_t seed)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    auto body = [&](uint64_t k1, uint64_t k2) {
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    };
    auto tail = [&](uint64_t k1, uint64_t k2, bool has2) {
        if (has2) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    };
    if constexpr (NW == 1) {
        tail(w[0], 0, false);
    } else if constexpr (NW == 2) {
        if (k == 16) body(w[0], w[1]); else tail(w[0], w[1], true);
    } else if constexpr (NW == 3) {
        body(w[0], w[1]); tail(w[2], 0, false);
    } else {
        body(w[0], w[1]);
        if (k == 32) body(w[2], w[3]); else tail(w[2], w[3], true);
    }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    return h1 + h2;
}

// upper-case 8 packed bytes: 'a'..'z' -> -32 (bytes >= 0x80 untouched, as the reference's signed char test)
__device__ __forceinline__ uint64_t upper8(uint64_t x)
{
    const uint64_t x7 = x & 0x7f7f7f7f7f7f7f7fULL;
    const uint64_t ge_a = x7 + 0x1f1f1f1f1f1f1f1fULL;   // bit7 set iff byte >= 0x61
    const uint64_t gt_z = x7 + 0x0505050505050505ULL;   // bit7 set iff byte >= 0x7b
    const uint64_t lower = ge_a & ~gt_z & ~x & 0x8080808080808080ULL;
    return x - (lower >> 2);
}

// compaction between steps: all threads call it (block-uniform); returns the new threshold
__device__ __noinline__ unsigned long long k1_compact_step(uint64_t *buf, uint32_t cap, uint32_t s, uint32_t *scratch,
                                                           uint32_t *s_count, unsigned long long *s_tau,
                                                           unsigned long long *g_tau)
{
    const uint32_t cnt = min(*s_count, cap);
    unsigned long long t2 = *s_tau;
    __syncthreads();
    const uint32_t m = cta_compact(buf, cnt, cap, s, scratch, &t2);
    if (threadIdx.x == 0) {
        *s_count = m;
        if (t2 < *s_tau) { *s_tau = t2; atomicMin(g_tau, t2); }
        const unsigned long long g = *reinterpret_cast<volatile unsigned long long *>(g_tau);
        if (g < *s_tau) *s_tau = g;
    }
    __syncthreads();
    return *s_tau;
}

template <int NW>
__global__ void __launch_bounds__(K1_THREADS) k1_sketch_segment_kernel(const K1Args a)
{
    extern __shared__ __align__(16) uint8_t k1_smem[];
    __shared__ uint32_t s_count, s_seg;
    __shared__ unsigned long long s_tau;
    __shared__ __align__(8) uint32_t s_scratch[40];
    uint64_t *buf = a.use_global ? a.gscratch + (size_t)blockIdx.x * a.cap : reinterpret_cast<uint64_t *>(k1_smem);
    const int k = a.k;
    const int last_shift = 56 - 8 * ((k - 1) & 7);                 // bit position of byte k-1 in its BE word
    const uint64_t last_mask = ~0ULL << last_shift;                 // keeps bytes 0..(k-1)%8 of the last word

    for (;;) {
        if (threadIdx.x == 0) s_seg = atomicAdd(a.seg_counter, 1u);
        __syncthreads();
        const uint32_t seg_id = s_seg;
        if (seg_id >= a.nseg) break;
        const K1Seg seg = a.segs[seg_id];
        const uint32_t s = a.sketch_size[seg.genome];
        if (threadIdx.x == 0) { s_count = 0; s_tau = a.tau[seg.genome]; }
        __syncthreads();

        // thread t owns windows [t*R, t*R + nwin_t); R multiple of K1_STEP
        uint32_t R = (seg.nwin + K1_THREADS - 1) / K1_THREADS;
        R = (R + K1_STEP - 1) / K1_STEP * K1_STEP;
        const uint32_t w0 = threadIdx.x * R;
        const uint32_t nwin_t = w0 < seg.nwin ? min(R, seg.nwin - w0) : 0u;
        const uint32_t nbases = R + (uint32_t)k - 1;               // bases every thread steps through (uniform)
        const unsigned long long p0 = seg.seq_start + w0;          // first base of this thread
        const uint64_t *wp = reinterpret_cast<const uint64_t *>(a.seq + (p0 & ~7ULL));
        const int mis = (int)(p0 & 7ULL) * 8;
        // does this thread have anything to read? (threads beyond the segment must not touch memory)
        const bool active = nwin_t > 0;
        uint64_t cur = active ? __ldg(wp) : 0;

        uint64_t F[NW], Rv[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) { F[i] = 0; Rv[i] = 0; }
        unsigned long long tau = s_tau;

        for (uint32_t b0 = 0; b0 < nbases; b0 += K1_STEP) {
            // next 8 bases of this thread: bytes [p0+b0, p0+b0+8) assembled from two aligned words
            const bool need = active && (b0 < nwin_t + (uint32_t)k - 1);
            const uint64_t nxt = need ? __ldg(wp + (b0 >> 3) + 1) : 0;
            uint64_t eight = mis ? ((cur >> mis) | (nxt << (64 - mis))) : cur;
            cur = nxt;
            eight = upper8(eight);
#pragma unroll
            for (int half = 0; half < K1_STEP / K1_VOTE; half++) {
#pragma unroll
                for (int jj = 0; jj < K1_VOTE; jj++) {
                    const int j = half * K1_VOTE + jj;
                    const uint32_t c = (uint32_t)(eight >> (8 * j)) & 0xffu;
                    const uint32_t comp = __byte_perm(0x43414754u, 0u, ((c >> 1) & 3u) | 0x4440u);   // "TGAC"[(c>>1)&3]
                    // forward strand: drop byte 0, append c at byte k-1
#pragma unroll
                    for (int i = 0; i < NW - 1; i++) F[i] = (F[i] << 8) | (F[i + 1] >> 56);
                    F[NW - 1] = (F[NW - 1] << 8) | ((uint64_t)c << last_shift);
                    // reverse-complement strand: prepend comp(c), drop byte k-1
#pragma unroll
                    for (int i = NW - 1; i > 0; i--) Rv[i] = (Rv[i] >> 8) | (Rv[i - 1] << 56);
                    Rv[0] = (Rv[0] >> 8) | ((uint64_t)comp << 56);
                    Rv[NW - 1] &= last_mask;
                    const uint32_t bi = b0 + j;                         // base index just consumed
                    if (bi >= (uint32_t)k - 1 && bi - ((uint32_t)k - 1) < nwin_t) {
                        bool fwd = true;                                // memcmp(fwd, rev, k) <= 0
#pragma unroll
                        for (int i = NW - 1; i >= 0; i--) fwd = (F[i] != Rv[i]) ? (F[i] < Rv[i]) : fwd;
                        uint64_t w[NW];
#pragma unroll
                        for (int i = 0; i < NW; i++) w[i] = bswap64(fwd ? F[i] : Rv[i]);
                        const uint64_t h = murmur_h1<NW>(w, k, a.seed);
                        if (h <= tau) {
                            const uint32_t pos = atomicAdd(&s_count, 1u);
                            if (pos < a.cap) buf[pos] = h;
                        }
                    }
                }
                // block-wide vote: compact when the appends of the next K1_VOTE bases could overflow.  A thread
                // may read s_count before slower threads finished this round, hence the 2x margin.
                const bool full = s_count + 2 * K1_STEP_CAND > a.cap;
                if (__syncthreads_or(full)) {
                    tau = k1_compact_step(buf, a.cap, s, s_scratch, &s_count, &s_tau, &a.tau[seg.genome]);
                }
            }
        }
        __syncthreads();
        {   // final compaction of this segment and write-out of the partial sketch
            const uint32_t cnt = min(s_count, a.cap);
            unsigned long long t2 = s_tau;
            const uint32_t m = cta_compact(buf, cnt, a.cap, s, s_scratch, &t2);
            if (threadIdx.x == 0) {
                a.partial_n[seg_id] = m;
                if (t2 < s_tau) atomicMin(&a.tau[seg.genome], t2);
            }
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) a.partial[seg.out_off + i] = buf[i];
        }
        __syncthreads();
    }
}

// merge the partial sketches of one genome: concat, sort, unique, truncate
struct K1MergeArgs {
    const uint64_t *partial;
    const uint32_t *partial_n;
    const K1Seg *segs;
    const uint32_t *genome_seg_begin;   // ngenomes+1
    const uint32_t *sketch_size;
    uint64_t *out; unsigned long long out_stride; uint32_t *out_n;
    uint32_t cap;                       // buffer capacity (power of two)
    uint64_t *gscratch; int use_global;
};

__global__ void __launch_bounds__(K1_THREADS) k1_merge_kernel(const K1MergeArgs a)
{
    extern __shared__ __align__(16) uint8_t k1_smem[];
    __shared__ __align__(8) uint32_t s_scratch[40];
    __shared__ uint32_t s_total;
    const uint32_t g = blockIdx.x;
    uint64_t *buf = a.use_global ? a.gscratch + (size_t)g * a.cap : reinterpret_cast<uint64_t *>(k1_smem);
    const uint32_t sb = a.genome_seg_begin[g], se = a.genome_seg_begin[g + 1];
    const uint32_t s = a.sketch_size[g];
    uint32_t total = 0;
    for (uint32_t sg = sb; sg < se; sg++) {
        const uint32_t m = a.partial_n[sg];
        for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) buf[total + i] = a.partial[a.segs[sg].out_off + i];
        total += m;
    }
    __syncthreads();
    unsigned long long dummy = 0;
    uint32_t m = total;
    if (se - sb > 1) m = cta_compact(buf, total, a.cap, s, s_scratch, &dummy);
    else if (m > s) m = s;
    if (threadIdx.x == 0) { a.out_n[g] = m; s_total = m; }
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) a.out[(size_t)g * a.out_stride + i] = buf[i];
}

uint32_t next_pow2(uint64_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }

template <int NW>
int launch_segments(tc_ctx *ctx, const K1Args &a, int grid, size_t smem)
{
    if (smem > 48 * 1024)
        TC_CUDA(cudaFuncSetAttribute(k1_sketch_segment_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1_sketch_segment_kernel<NW><<<grid, K1_THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    return 0;
}

// seq_dev: sequence bytes on the device (readable 16 B past the end). Everything else host.
int sketch_minhash_impl(tc_ctx *ctx, const uint8_t *seq_dev, const uint64_t *rec_off, const uint32_t *rec2genome,
                        uint64_t nrec, uint32_t ngenomes, int k, const uint32_t *sketch_size, uint32_t seed,
                        uint64_t *out_dev, uint64_t out_stride, uint32_t *out_n_dev)
{
    TC_CHECK(k >= 1 && k <= 32, "tc_sketch_minhash: k=%d unsupported (1..32, as the reference's AVX2 path)", k);
    TC_CHECK(ngenomes > 0, "tc_sketch_minhash: no genomes");
    uint32_t smax = 0;
    for (uint32_t g = 0; g < ngenomes; g++) {
        TC_CHECK(sketch_size[g] > 0, "tc_sketch_minhash: sketch_size[%u] == 0", g);
        smax = std::max(smax, sketch_size[g]);
    }
    TC_CHECK(out_stride >= smax, "tc_sketch_minhash: out_stride %llu < max sketch size %u",
             (unsigned long long)out_stride, smax);
    // ---- segments
    uint64_t total_win = 0;
    for (uint64_t r = 0; r < nrec; r++) {
        TC_CHECK(rec_off[r + 1] >= rec_off[r], "tc_sketch_minhash: record offsets not monotone");
        TC_CHECK(r == 0 || rec2genome[r] >= rec2genome[r - 1], "tc_sketch_minhash: rec2genome must be non-decreasing");
        TC_CHECK(rec2genome[r] < ngenomes, "tc_sketch_minhash: rec2genome[%llu] out of range", (unsigned long long)r);
        const uint64_t len = rec_off[r + 1] - rec_off[r];
        if (len >= (uint64_t)k) total_win += len - k + 1;
    }
    const uint64_t want_ctas = (uint64_t)ctx->num_sms * 24;
    uint64_t segw = std::min<uint64_t>(std::max<uint64_t>(total_win / std::max<uint64_t>(want_ctas, 1), 1u << 15), 1u << 20);
    segw = (segw + K1_THREADS * K1_STEP - 1) / (K1_THREADS * K1_STEP) * (K1_THREADS * K1_STEP);
    std::vector<K1Seg> segs;
    std::vector<uint32_t> gbegin(ngenomes + 1, 0);
    uint64_t out_off = 0;
    {
        uint64_t r = 0;
        for (uint32_t g = 0; g < ngenomes; g++) {
            gbegin[g] = (uint32_t)segs.size();
            for (; r < nrec && rec2genome[r] == g; r++) {
                const uint64_t len = rec_off[r + 1] - rec_off[r];
                if (len < (uint64_t)k) continue;
                const uint64_t nwin = len - k + 1;
                for (uint64_t w = 0; w < nwin; w += segw) {
                    K1Seg s;
                    s.seq_start = rec_off[r] + w;
                    s.nwin = (uint32_t)std::min<uint64_t>(segw, nwin - w);
                    s.genome = g;
                    s.out_off = out_off;
                    out_off += sketch_size[g];
                    segs.push_back(s);
                }
            }
        }
        gbegin[ngenomes] = (uint32_t)segs.size();
    }
    const uint32_t nseg = (uint32_t)segs.size();
    uint32_t max_seg_per_genome = 1;
    uint64_t max_merge = 1;
    for (uint32_t g = 0; g < ngenomes; g++) {
        max_seg_per_genome = std::max(max_seg_per_genome, gbegin[g + 1] - gbegin[g]);
        max_merge = std::max<uint64_t>(max_merge, (uint64_t)(gbegin[g + 1] - gbegin[g]) * sketch_size[g]);
    }
    cudaStream_t st = ctx->stream;
    // ---- device metadata
    K1Seg *d_segs = nullptr; uint32_t *d_gbegin = nullptr, *d_ss = nullptr, *d_pn = nullptr;
    unsigned long long *d_tau = nullptr; uint64_t *d_partial = nullptr; unsigned int *d_counter = nullptr;
    uint64_t *d_scr1 = nullptr, *d_scr2 = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(d_segs); cudaFree(d_gbegin); cudaFree(d_ss); cudaFree(d_pn); cudaFree(d_tau); cudaFree(d_partial);
        cudaFree(d_counter); cudaFree(d_scr1); cudaFree(d_scr2);
    };
#define K1_TRY(x) do { if ((x) != cudaSuccess) { tc_set_error("tc_sketch_minhash: %s failed: %s", #x, cudaGetErrorString(cudaGetLastError())); cleanup(); return 1; } } while (0)
    K1_TRY(cudaMalloc(&d_segs, std::max<size_t>(nseg, 1) * sizeof(K1Seg)));
    K1_TRY(cudaMalloc(&d_gbegin, (size_t)(ngenomes + 1) * 4));
    K1_TRY(cudaMalloc(&d_ss, (size_t)ngenomes * 4));
    K1_TRY(cudaMalloc(&d_pn, std::max<size_t>(nseg, 1) * 4));
    K1_TRY(cudaMalloc(&d_tau, (size_t)ngenomes * 8));
    K1_TRY(cudaMalloc(&d_partial, std::max<uint64_t>(out_off, 1) * 8));
    K1_TRY(cudaMalloc(&d_counter, 4));
    K1_TRY(cudaMemcpyAsync(d_segs, segs.data(), nseg * sizeof(K1Seg), cudaMemcpyHostToDevice, st));
    K1_TRY(cudaMemcpyAsync(d_gbegin, gbegin.data(), (size_t)(ngenomes + 1) * 4, cudaMemcpyHostToDevice, st));
    K1_TRY(cudaMemcpyAsync(d_ss, sketch_size, (size_t)ngenomes * 4, cudaMemcpyHostToDevice, st));
    K1_TRY(cudaMemsetAsync(d_tau, 0xFF, (size_t)ngenomes * 8, st));
    K1_TRY(cudaMemsetAsync(d_counter, 0, 4, st));
    K1_TRY(cudaMemsetAsync(d_pn, 0, std::max<size_t>(nseg, 1) * 4, st));

    if (nseg > 0) {
        K1Args a{};
        a.seq = seq_dev; a.segs = d_segs; a.nseg = nseg; a.k = k; a.seed = seed; a.sketch_size = d_ss; a.tau = d_tau;
        a.partial = d_partial; a.partial_n = d_pn; a.seg_counter = d_counter;
        a.cap = std::max<uint32_t>(next_pow2((uint64_t)smax + 2 * K1_STEP_CAND + 1), 4096);
        size_t smem = (size_t)a.cap * 8;
        a.use_global = smem > 160 * 1024;
        int per_sm = a.use_global ? 4 : (int)std::max<size_t>(1, std::min<size_t>(6, (200 * 1024) / (smem + 1024)));
        int grid = (int)std::min<uint64_t>(nseg, (uint64_t)ctx->num_sms * per_sm);
        if (a.use_global) {
            K1_TRY(cudaMalloc(&d_scr1, (size_t)grid * a.cap * 8));
            a.gscratch = d_scr1;
            smem = 0;
        }
        const int nw = (k + 7) / 8;
        rc = nw == 1 ? launch_segments<1>(ctx, a, grid, smem) : nw == 2 ? launch_segments<2>(ctx, a, grid, smem)
           : nw == 3 ? launch_segments<3>(ctx, a, grid, smem) : launch_segments<4>(ctx, a, grid, smem);
        if (rc) { cleanup(); return rc; }
    }
    {
        K1MergeArgs m{};
        m.partial = d_partial; m.partial_n = d_pn; m.segs = d_segs; m.genome_seg_begin = d_gbegin; m.sketch_size = d_ss;
        m.out = out_dev; m.out_stride = out_stride; m.out_n = out_n_dev;
        m.cap = std::max<uint32_t>(next_pow2(max_merge), 2);
        size_t smem = (size_t)m.cap * 8;
        m.use_global = smem > 160 * 1024;
        if (m.use_global) {
            K1_TRY(cudaMalloc(&d_scr2, (size_t)ngenomes * m.cap * 8));
            m.gscratch = d_scr2;
            smem = 0;
        }
        if (smem > 48 * 1024)
            K1_TRY(cudaFuncSetAttribute(k1_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1_merge_kernel<<<ngenomes, K1_THREADS, smem, st>>>(m);
        ctx->launches++;
        K1_TRY(cudaGetLastError());
    }
    K1_TRY(cudaStreamSynchronize(st));
#undef K1_TRY
    cleanup();
    return 0;
}

}  // namespace

extern "C" int tc_sketch_minhash_dev(tc_ctx *ctx, const uint8_t *seqs_dev, const uint64_t *rec_off,
                                     const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes, int k,
                                     const uint32_t *sketch_size, uint32_t seed, uint64_t *out_hashes_dev,
                                     uint64_t out_stride, uint32_t *out_n_dev)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    return sketch_minhash_impl(ctx, seqs_dev, rec_off, rec2genome, nrec, ngenomes, k, sketch_size, seed,
                               out_hashes_dev, out_stride, out_n_dev);
}

extern "C" int tc_sketch_minhash(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off, const uint32_t *rec2genome,
                                 uint64_t nrec, uint32_t ngenomes, int k, const uint32_t *sketch_size, uint32_t seed,
                                 uint64_t *out_hashes, uint64_t out_stride, uint32_t *out_n)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    const uint64_t nbytes = nrec ? rec_off[nrec] : 0;
    uint8_t *d_seq = nullptr; uint64_t *d_out = nullptr; uint32_t *d_n = nullptr;
    TC_CUDA(cudaMalloc(&d_seq, nbytes + 64));
    int rc = 1;
    do {
        if (cudaMalloc(&d_out, std::max<uint64_t>((uint64_t)ngenomes * out_stride, 1) * 8) != cudaSuccess ||
            cudaMalloc(&d_n, std::max<uint32_t>(ngenomes, 1) * 4) != cudaSuccess) {
            tc_set_error("tc_sketch_minhash: out of device memory");
            break;
        }
        if (cudaMemcpyAsync(d_seq, seqs, nbytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemsetAsync(d_seq + nbytes, 0, 64, ctx->stream) != cudaSuccess) {
            tc_set_error("tc_sketch_minhash: H2D copy failed");
            break;
        }
        rc = sketch_minhash_impl(ctx, d_seq, rec_off, rec2genome, nrec, ngenomes, k, sketch_size, seed, d_out,
                                 out_stride, d_n);
        if (rc) break;
        if (cudaMemcpyAsync(out_hashes, d_out, (uint64_t)ngenomes * out_stride * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(out_n, d_n, (size_t)ngenomes * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            tc_set_error("tc_sketch_minhash: D2H copy failed");
            rc = 1;
        }
    } while (0);
    cudaFree(d_seq); cudaFree(d_out); cudaFree(d_n);
    return rc;
}
