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
// Design (DESIGN.md "K1"): the path is integer-issue bound (Murmur3 = ten 64-bit multiplies per k-mer), not
// HBM bound.  A genome is cut into segments of whole windows; one CTA sketches one segment: every thread
// walks its own contiguous run of windows with the next 8*(NW+1) bases held in registers twice -- as read
// (forward strand, already the little-endian word layout Murmur3 consumes) and complemented + shifted so
// that the reverse-complement k-mer, read backwards, is a byte range whose words compare like memcmp (see
// the comment on k1_sketch_segment_kernel).  There is no special path for non-ACGT bytes.  Hashes not
// above the running threshold tau (the s-th smallest distinct hash seen so far) are appended to a
// shared-memory candidate buffer which is sorted/deduplicated/truncated (bitonic) when it fills; tau is
// shared between the segments of a genome through a global atomicMin.  A second kernel merges the
// per-segment partial sketches of each genome (prefixes cut at tau, compaction on overflow).  The host
// entry point pipelines batches of genomes over two staging slots (H2D of batch c+1 under the kernels of c).
#include "tc_common.cuh"
#include "k_sort.cuh"
#include "k_stage.cuh"
#include <algorithm>
#include <thread>
#include <cstdlib>
#include <cstring>

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

// complement of 4 packed bytes: "TGAC"[(c>>1)&3] per byte (MinHash.cpp:1185-1195), two PRMT table look-ups
__device__ __forceinline__ uint32_t comp4(uint32_t v)
{
    const uint32_t idx = (v >> 1) & 0x03030303u;           // byte t = 2-bit table index of base t
    const uint32_t t = idx | (idx >> 4);                    // byte0 = i0 | i1<<4, byte2 = i2 | i3<<4
    const uint32_t sel = __byte_perm(t, 0u, 0x4420);        // selector nibbles i0, i1, i2, i3
    return __byte_perm(0x43414754u, 0u, sel);               // table bytes 'T','G','A','C'
}
__device__ __forceinline__ uint64_t comp8(uint64_t x)
{
    return ((uint64_t)comp4((uint32_t)(x >> 32)) << 32) | comp4((uint32_t)x);
}
// bytes [8i+u, 8i+u+8) of a little-endian word array (u compile-time after unrolling)
template <int N>
__device__ __forceinline__ uint64_t ext_word(const uint64_t (&b)[N], int i, int u)
{
    return u == 0 ? b[i] : (b[i] >> (8 * u)) | (b[i + 1] << (64 - 8 * u));
}

// Inner loop (second generation).  A thread keeps the next 8*(NW+1) bases of its run in registers twice:
//   fb[]  the upper-cased bytes in memory order (little-endian words): the forward k-mer of the window at
//         byte offset u is the byte range [u, u+k) -- already the word layout Murmur3 reads;
//   cb[]  the complemented bytes, stored shifted by r-8 bytes (r = k - 8*(NW-1) = bytes in the last k-mer
//         word), so that the reverse-complement k-mer, read BACKWARDS, is the byte range [u, u+8*NW) of cb:
//         word NW-1-i of that range is word i of the reverse-complement strand in BIG-endian byte order.
// Big-endian words compare like memcmp, so the canonical-strand test is one 64-bit compare of
// bswap(forward word 0) with the top cb word (later words only when the first 8 bytes tie); the words of the
// chosen strand are produced by one PRMT per 32-bit half (identity on the forward half or byte-reversal of
// the cb half, selected by the permute control).  All byte offsets inside the 8-window step are
// compile-time constants; k enters only through two masks and one funnel shift per 8 loaded bases.
template <int NW>
__global__ void __launch_bounds__(K1_THREADS) k1_sketch_segment_kernel(const K1Args a)
{
    constexpr int NB = NW + 1;
    extern __shared__ __align__(16) uint8_t k1_smem[];
    __shared__ uint32_t s_count, s_seg;
    __shared__ unsigned long long s_tau;
    __shared__ __align__(8) uint32_t s_scratch[40];
    uint64_t *buf = a.use_global ? a.gscratch + (size_t)blockIdx.x * a.cap : reinterpret_cast<uint64_t *>(k1_smem);
    const int k = a.k;
    const int r = k - 8 * (NW - 1);                                  // 1..8 bytes in the last k-mer word
    const uint64_t lomask = r == 8 ? ~0ULL : ((1ULL << (8 * r)) - 1ULL);   // forward last word: low r bytes
    const int csh = 8 * (8 - r);                                     // cb is stored csh bits "late"
    const uint64_t himask = ~0ULL << csh;                            // reverse last word: high r bytes

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
        const uint32_t nbytes_t = nwin_t ? nwin_t + (uint32_t)k - 1 : 0u;   // bytes this thread may touch
        const unsigned long long p0 = seg.seq_start + w0;                    // first base of this thread
        const uint64_t *wp = reinterpret_cast<const uint64_t *>(a.seq + (p0 & ~7ULL));
        const int mis = (int)(p0 & 7ULL) * 8;
        uint64_t cur = nbytes_t ? __ldg(wp) : 0;
        uint32_t jw = 0;                                                     // next 8-byte word of the run
        // bytes [p0+8*jw, +8), upper-cased; zero when the thread needs none of them
        auto next_word = [&]() -> uint64_t {
            const bool need = 8u * jw < nbytes_t;
            const uint64_t nxt = need ? __ldg(wp + jw + 1) : 0;
            const uint64_t w = mis ? ((cur >> mis) | (nxt << (64 - mis))) : cur;
            cur = nxt;
            jw++;
            return need ? upper8(w) : 0;
        };

        uint64_t fb[NB], cb[NB];
        uint64_t cprev = 0;
#pragma unroll
        for (int i = 0; i < NB; i++) {
            fb[i] = next_word();
            const uint64_t c = comp8(fb[i]);
            cb[i] = (c << csh) | ((cprev >> 1) >> (63 - csh));
            cprev = c;
        }
        unsigned long long tau = s_tau;

        for (uint32_t g0 = 0; g0 < R; g0 += K1_STEP) {
#pragma unroll
            for (int half = 0; half < K1_STEP / K1_VOTE; half++) {
#pragma unroll
                for (int jj = 0; jj < K1_VOTE; jj++) {
                    const int u = half * K1_VOTE + jj;
                    if (g0 + u < nwin_t) {
                        uint64_t fw[NW], rx[NW];
#pragma unroll
                        for (int i = 0; i < NW; i++) fw[i] = ext_word<NB>(fb, i, u);
                        fw[NW - 1] &= lomask;
#pragma unroll
                        for (int i = 0; i < NW - 1; i++) rx[i] = ext_word<NB>(cb, NW - 1 - i, u);
                        rx[NW - 1] = ext_word<NB>(cb, 0, u) & himask;
                        // memcmp(forward, reverse, k) <= 0 on big-endian words
                        bool fwd = true;
                        {
                            const uint64_t f0 = bswap64(fw[0]);
                            if (f0 != rx[0]) fwd = f0 < rx[0];
                            else {
#pragma unroll
                                for (int i = NW - 1; i >= 1; i--) {
                                    const uint64_t fi = bswap64(fw[i]);
                                    fwd = (fi != rx[i]) ? (fi < rx[i]) : fwd;
                                }
                            }
                        }
                        const uint32_t sel = fwd ? 0x3210u : 0x4567u;   // identity on a / byte reversal of b
                        uint64_t w[NW];
#pragma unroll
                        for (int i = 0; i < NW; i++) {
                            const uint32_t lo = __byte_perm((uint32_t)fw[i], (uint32_t)(rx[i] >> 32), sel);
                            const uint32_t hi = __byte_perm((uint32_t)(fw[i] >> 32), (uint32_t)rx[i], sel);
                            w[i] = ((uint64_t)hi << 32) | lo;
                        }
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
            // slide the register window by one word (8 bases)
#pragma unroll
            for (int i = 0; i < NB - 1; i++) { fb[i] = fb[i + 1]; cb[i] = cb[i + 1]; }
            fb[NB - 1] = next_word();
            const uint64_t c = comp8(fb[NB - 1]);
            cb[NB - 1] = (c << csh) | ((cprev >> 1) >> (63 - csh));
            cprev = c;
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

// merge the partial sketches of one genome.  The partials are sorted and distinct; they are appended to the
// buffer segment by segment, cut at the running threshold tau (the s-th smallest distinct value merged so far;
// binary search on the sorted partial), and the buffer is sorted/deduplicated/truncated whenever the next prefix
// does not fit.  tau tightens after every compaction, so a genome with hundreds of segments needs only a few of
// them; repeated k-mers shared by many segments just cost more compactions.
struct K1MergeArgs {
    const uint64_t *partial;
    const uint32_t *partial_n;
    const K1Seg *segs;
    const uint32_t *genome_seg_begin;   // ngenomes+1
    const uint32_t *sketch_size;
    uint64_t *out; unsigned long long out_stride; uint32_t *out_n;
    uint32_t cap;                       // buffer capacity (power of two, >= 2 * largest sketch size)
    uint64_t *gscratch; int use_global;
};

__global__ void __launch_bounds__(K1_THREADS) k1_merge_kernel(const K1MergeArgs a)
{
    extern __shared__ __align__(16) uint8_t k1_smem[];
    __shared__ __align__(8) uint32_t s_scratch[40];
    const uint32_t g = blockIdx.x;
    uint64_t *buf = a.use_global ? a.gscratch + (size_t)g * a.cap : reinterpret_cast<uint64_t *>(k1_smem);
    const uint32_t sb = a.genome_seg_begin[g], se = a.genome_seg_begin[g + 1];
    const uint32_t s = a.sketch_size[g];
    uint32_t count = 0, nsrc = 0;                 // block-uniform
    unsigned long long tau = ~0ULL;
    for (uint32_t sg = sb; sg < se; sg++) {
        const uint32_t n = a.partial_n[sg];
        const uint64_t *p = a.partial + a.segs[sg].out_off;
        auto prefix_len = [&]() {                 // #values <= tau (every thread runs the same search)
            uint32_t lo = 0, hi = n;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (p[mid] <= tau) lo = mid + 1; else hi = mid;
            }
            return lo;
        };
        uint32_t m = prefix_len();
        if (m == 0) continue;
        if (count + m > a.cap) {
            __syncthreads();
            count = cta_compact(buf, count, a.cap, s, s_scratch, &tau);
            __syncthreads();
            m = prefix_len();
            if (m == 0) continue;
        }
        for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) buf[count + i] = p[i];
        count += m;
        nsrc++;
    }
    __syncthreads();
    if (nsrc > 1) count = cta_compact(buf, count, a.cap, s, s_scratch, &tau);
    else if (count > s) count = s;
    if (threadIdx.x == 0) a.out_n[g] = count;
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) a.out[(size_t)g * a.out_stride + i] = buf[i];
}

uint32_t next_pow2(uint64_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }

template <int NW>
int launch_segments(tc_ctx *ctx, cudaStream_t st, const K1Args &a, int grid, size_t smem)
{
    if (smem > 48 * 1024)
        TC_CUDA(cudaFuncSetAttribute(k1_sketch_segment_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1_sketch_segment_kernel<NW><<<grid, K1_THREADS, smem, st>>>(a);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------ host driver
// A batch (whole genomes) is planned on the host (segments), its metadata goes through a pinned staging
// buffer, and two kernels are enqueued.  Nothing here synchronises: the host-buffer entry point pipelines
// batches over two slots so that the H2D copy of batch c+1 (copy stream) overlaps the kernels of batch c.
struct K1Slot {
    uint8_t *d_seq = nullptr;  size_t seq_cap = 0;      // staged sequence bytes (host-buffer path only)
    uint8_t *h_meta = nullptr; size_t h_meta_cap = 0;   // pinned: segments, genome->segment ranges, sketch sizes
    uint8_t *d_meta = nullptr; size_t d_meta_cap = 0;   // the same on the device + partial sketches, counters
    cudaEvent_t copied = nullptr, done = nullptr;
    bool used = false;
    TcPackBuf pack;                                      // packed 2-bit transport (k_stage.cuh)
};
struct K1Workspace {
    K1Slot slot[2];
    cudaStream_t copy_stream = nullptr;
    uint64_t *d_out = nullptr; size_t out_cap = 0;      // staged outputs (host-buffer path)
    uint32_t *d_n = nullptr;   size_t n_cap = 0;
};

K1Workspace *k1_ws(tc_ctx *ctx)
{
    if (!ctx->k1_ws) {
        K1Workspace *w = new K1Workspace();
        if (cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete w; return nullptr; }
        for (K1Slot &sl : w->slot) {
            cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
        }
        ctx->k1_ws = w;
    }
    return static_cast<K1Workspace *>(ctx->k1_ws);
}

struct K1Plan {
    std::vector<K1Seg> segs;
    std::vector<uint32_t> gbegin;
    uint64_t out_off = 0, max_merge = 1;
    uint32_t smax = 0;
};

inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

// records [r0, r1) = genomes [g0, g1); sequence byte x of the input lives at seq_dev[x - base_off]
void k1_plan(const tc_ctx *ctx, const uint64_t *rec_off, const uint32_t *rec2genome, uint64_t r0, uint64_t r1,
             uint32_t g0, uint32_t g1, int k, const uint32_t *sketch_size, uint64_t base_off, K1Plan &pl)
{
    const uint32_t ngl = g1 - g0;
    uint64_t total_win = 0;
    for (uint64_t r = r0; r < r1; r++) {
        const uint64_t len = rec_off[r + 1] - rec_off[r];
        if (len >= (uint64_t)k) total_win += len - k + 1;
    }
    // segments of 128 Ki .. 1 Mi windows: a segment re-discovers the bottom-s set on its own (about
    // s*(1+ln(windows/s)) candidates pass its filter), so shorter segments spend more time sorting than hashing
    // one wave of resident CTAs (6 per SM at s = 1000) when the batch is small enough for that
    const uint64_t want_ctas = (uint64_t)ctx->num_sms * 6;
    uint64_t segw = std::min<uint64_t>(std::max<uint64_t>(total_win / std::max<uint64_t>(want_ctas, 1), 1u << 17), 1u << 20);
    segw = (segw + K1_THREADS * K1_STEP - 1) / (K1_THREADS * K1_STEP) * (K1_THREADS * K1_STEP);
    pl.segs.clear();
    pl.gbegin.assign(ngl + 1, 0);
    pl.out_off = 0; pl.max_merge = 1; pl.smax = 0;
    uint64_t r = r0;
    for (uint32_t g = g0; g < g1; g++) {
        pl.gbegin[g - g0] = (uint32_t)pl.segs.size();
        pl.smax = std::max(pl.smax, sketch_size[g]);
        for (; r < r1 && rec2genome[r] == g; r++) {
            const uint64_t len = rec_off[r + 1] - rec_off[r];
            if (len < (uint64_t)k) continue;
            const uint64_t nwin = len - k + 1;
            for (uint64_t w = 0; w < nwin; w += segw) {
                K1Seg sg;
                sg.seq_start = rec_off[r] - base_off + w;
                sg.nwin = (uint32_t)std::min<uint64_t>(segw, nwin - w);
                sg.genome = g - g0;
                sg.out_off = pl.out_off;
                pl.out_off += sketch_size[g];
                pl.segs.push_back(sg);
            }
        }
        const uint32_t nsg = (uint32_t)pl.segs.size() - pl.gbegin[g - g0];
        pl.max_merge = std::max<uint64_t>(pl.max_merge, (uint64_t)nsg * sketch_size[g]);
    }
    pl.gbegin[ngl] = (uint32_t)pl.segs.size();
}

// enqueue the two kernels of one planned batch on `st`; out_dev / out_n_dev / sketch_size point at genome g0
int k1_enqueue(tc_ctx *ctx, K1Slot &sl, cudaStream_t st, const uint8_t *seq_dev, const K1Plan &pl, uint32_t ngl, int k,
               uint32_t seed, const uint32_t *sketch_size, uint64_t *out_dev, uint64_t out_stride, uint32_t *out_n_dev)
{
    const uint32_t nseg = (uint32_t)pl.segs.size();
    // ---- layout of the metadata block
    const size_t o_segs = 0;
    const size_t o_gbegin = al16(o_segs + std::max<size_t>(nseg, 1) * sizeof(K1Seg));
    const size_t o_ss = al16(o_gbegin + (size_t)(ngl + 1) * 4);
    const size_t host_bytes = al16(o_ss + (size_t)ngl * 4);           // uploaded part
    const size_t o_pn = host_bytes;
    const size_t o_tau = al16(o_pn + std::max<size_t>(nseg, 1) * 4);
    const size_t o_counter = al16(o_tau + (size_t)ngl * 8);
    const size_t o_partial = al16(o_counter + 16);
    size_t total = al16(o_partial + std::max<uint64_t>(pl.out_off, 1) * 8);

    K1Args a{};
    a.cap = std::max<uint32_t>(next_pow2((uint64_t)pl.smax + 2 * K1_STEP_CAND + 1), 4096);
    size_t smem1 = (size_t)a.cap * 8;
    a.use_global = smem1 > 160 * 1024;
    const int per_sm = a.use_global ? 4 : (int)std::max<size_t>(1, std::min<size_t>(6, (200 * 1024) / (smem1 + 1024)));
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(nseg, (uint64_t)ctx->num_sms * per_sm));
    size_t o_scr1 = 0, o_scr2 = 0;
    if (a.use_global) { o_scr1 = total; total = al16(total + (size_t)grid * a.cap * 8); smem1 = 0; }
    K1MergeArgs m{};
    m.cap = std::max<uint32_t>(next_pow2(2ull * pl.smax), 4096);
    size_t smem2 = (size_t)m.cap * 8;
    m.use_global = smem2 > 160 * 1024;
    if (m.use_global) { o_scr2 = total; total = al16(total + (size_t)ngl * m.cap * 8); smem2 = 0; }

    // ---- capacity (grows only; the caller guarantees the slot is idle)
    if (host_bytes > sl.h_meta_cap) {
        if (sl.h_meta) cudaFreeHost(sl.h_meta);
        sl.h_meta = nullptr; sl.h_meta_cap = 0;
        TC_CUDA(cudaMallocHost(&sl.h_meta, host_bytes * 2));
        sl.h_meta_cap = host_bytes * 2;
    }
    if (total > sl.d_meta_cap) {
        if (sl.d_meta) cudaFree(sl.d_meta);
        sl.d_meta = nullptr; sl.d_meta_cap = 0;
        TC_CUDA(cudaMalloc(&sl.d_meta, total + total / 4));
        sl.d_meta_cap = total + total / 4;
    }
    if (nseg) memcpy(sl.h_meta + o_segs, pl.segs.data(), (size_t)nseg * sizeof(K1Seg));
    memcpy(sl.h_meta + o_gbegin, pl.gbegin.data(), (size_t)(ngl + 1) * 4);
    memcpy(sl.h_meta + o_ss, sketch_size, (size_t)ngl * 4);
    TC_CUDA(cudaMemcpyAsync(sl.d_meta, sl.h_meta, host_bytes, cudaMemcpyHostToDevice, st));
    TC_CUDA(cudaMemsetAsync(sl.d_meta + o_pn, 0, o_tau - o_pn, st));
    TC_CUDA(cudaMemsetAsync(sl.d_meta + o_tau, 0xFF, (size_t)ngl * 8, st));
    TC_CUDA(cudaMemsetAsync(sl.d_meta + o_counter, 0, 16, st));

    const K1Seg *d_segs = reinterpret_cast<const K1Seg *>(sl.d_meta + o_segs);
    const uint32_t *d_ss = reinterpret_cast<const uint32_t *>(sl.d_meta + o_ss);
    uint32_t *d_pn = reinterpret_cast<uint32_t *>(sl.d_meta + o_pn);
    uint64_t *d_partial = reinterpret_cast<uint64_t *>(sl.d_meta + o_partial);
    if (nseg > 0) {
        a.seq = seq_dev; a.segs = d_segs; a.nseg = nseg; a.k = k; a.seed = seed; a.sketch_size = d_ss;
        a.tau = reinterpret_cast<unsigned long long *>(sl.d_meta + o_tau);
        a.partial = d_partial; a.partial_n = d_pn;
        a.seg_counter = reinterpret_cast<unsigned int *>(sl.d_meta + o_counter);
        a.gscratch = a.use_global ? reinterpret_cast<uint64_t *>(sl.d_meta + o_scr1) : nullptr;
        const int nw = (k + 7) / 8;
        const int rc = nw == 1 ? launch_segments<1>(ctx, st, a, grid, smem1) : nw == 2 ? launch_segments<2>(ctx, st, a, grid, smem1)
                     : nw == 3 ? launch_segments<3>(ctx, st, a, grid, smem1) : launch_segments<4>(ctx, st, a, grid, smem1);
        if (rc) return rc;
    }
    m.partial = d_partial; m.partial_n = d_pn; m.segs = d_segs;
    m.genome_seg_begin = reinterpret_cast<const uint32_t *>(sl.d_meta + o_gbegin); m.sketch_size = d_ss;
    m.out = out_dev; m.out_stride = out_stride; m.out_n = out_n_dev;
    m.gscratch = m.use_global ? reinterpret_cast<uint64_t *>(sl.d_meta + o_scr2) : nullptr;
    if (smem2 > 48 * 1024)
        TC_CUDA(cudaFuncSetAttribute(k1_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    k1_merge_kernel<<<ngl, K1_THREADS, smem2, st>>>(m);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    sl.used = true;
    return 0;
}

int k1_validate(const uint64_t *rec_off, const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes, int k,
                const uint32_t *sketch_size, uint64_t out_stride)
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
    for (uint64_t r = 0; r < nrec; r++) {
        TC_CHECK(rec_off[r + 1] >= rec_off[r], "tc_sketch_minhash: record offsets not monotone");
        TC_CHECK(r == 0 || rec2genome[r] >= rec2genome[r - 1], "tc_sketch_minhash: rec2genome must be non-decreasing");
        TC_CHECK(rec2genome[r] < ngenomes, "tc_sketch_minhash: rec2genome[%llu] out of range", (unsigned long long)r);
    }
    return 0;
}

}  // namespace

void tc_k1_free(tc_ctx *ctx)
{
    K1Workspace *w = static_cast<K1Workspace *>(ctx->k1_ws);
    if (!w) return;
    for (K1Slot &sl : w->slot) {
        cudaFree(sl.d_seq); cudaFree(sl.d_meta);
        sl.pack.release();
        if (sl.h_meta) cudaFreeHost(sl.h_meta);
        if (sl.copied) cudaEventDestroy(sl.copied);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    cudaFree(w->d_out); cudaFree(w->d_n);
    if (w->copy_stream) cudaStreamDestroy(w->copy_stream);
    delete w;
    ctx->k1_ws = nullptr;
}

extern "C" void *tc_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        tc_set_error("tc_host_alloc: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
extern "C" void tc_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" uint64_t tc_last_h2d_bytes(const tc_ctx *ctx) { return ctx ? ctx->last_h2d_bytes : 0; }

extern "C" int tc_sketch_minhash_dev(tc_ctx *ctx, const uint8_t *seqs_dev, const uint64_t *rec_off,
                                     const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes, int k,
                                     const uint32_t *sketch_size, uint32_t seed, uint64_t *out_hashes_dev,
                                     uint64_t out_stride, uint32_t *out_n_dev)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    if (k1_validate(rec_off, rec2genome, nrec, ngenomes, k, sketch_size, out_stride)) return 1;
    K1Workspace *ws = k1_ws(ctx);
    TC_CHECK(ws, "tc_sketch_minhash: cannot create the workspace");
    K1Plan pl;
    k1_plan(ctx, rec_off, rec2genome, 0, nrec, 0, ngenomes, k, sketch_size, 0, pl);
    if (k1_enqueue(ctx, ws->slot[0], ctx->stream, seqs_dev, pl, ngenomes, k, seed, sketch_size, out_hashes_dev,
                   out_stride, out_n_dev)) return 1;
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tc_sketch_minhash(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off, const uint32_t *rec2genome,
                                 uint64_t nrec, uint32_t ngenomes, int k, const uint32_t *sketch_size, uint32_t seed,
                                 uint64_t *out_hashes, uint64_t out_stride, uint32_t *out_n)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    if (k1_validate(rec_off, rec2genome, nrec, ngenomes, k, sketch_size, out_stride)) return 1;
    K1Workspace *ws = k1_ws(ctx);
    TC_CHECK(ws, "tc_sketch_minhash: cannot create the workspace");
    cudaStream_t st = ctx->stream, cs = ws->copy_stream;
    // staged outputs
    const size_t out_elems = std::max<uint64_t>((uint64_t)ngenomes * out_stride, 1);
    if (out_elems > ws->out_cap) {
        cudaFree(ws->d_out); ws->d_out = nullptr; ws->out_cap = 0;
        TC_CUDA(cudaMalloc(&ws->d_out, out_elems * 8));
        ws->out_cap = out_elems;
    }
    if (ngenomes > ws->n_cap) {
        cudaFree(ws->d_n); ws->d_n = nullptr; ws->n_cap = 0;
        TC_CUDA(cudaMalloc(&ws->d_n, (size_t)ngenomes * 4));
        ws->n_cap = ngenomes;
    }
    // ---- batches of whole genomes, about TCLUST_SKETCH_CHUNK_MB of sequence each.  Default: a quarter of the call's
    // sequence, between 128 MB and 1 GB -- the copy of batch c+1 hides behind the kernels of batch c only if those are
    // the shorter of the two, and the segment kernel is efficient from ~100 five-Mbp genomes per launch up (every
    // segment pays its own bottom-s warm-up: 25 genomes per launch run at a third of the full-size rate; measured,
    // profiles/r1m_sketch_chunk_sweep.txt)
    const uint64_t seq_bytes = nrec ? rec_off[nrec] : 0;
    const int pack_threads = tc_pack_threads();
    // (with the packed transport the host packer of batch c+1 runs beside the GPU work of batch c: a sixth of the call,
    //  measured 42.2 ms per 2.56 GB at 448 MB batches against 44.5 at 640 and 47.8 at 768)
    uint64_t chunk_bytes = std::min<uint64_t>(1ull << 30, std::max<uint64_t>(128ull << 20, seq_bytes / (pack_threads > 0 ? 6 : 4)));
    if (const char *e = getenv("TCLUST_SKETCH_CHUNK_MB")) chunk_bytes = std::max<uint64_t>(strtoull(e, nullptr, 10), 1) << 20;
    ctx->last_h2d_bytes = 0;
    K1Plan pl;
    uint64_t r0 = 0;
    uint32_t g0 = 0, batch = 0;
    int rc = 0;
    while (g0 < ngenomes && !rc) {
        // records of genomes [g0, g1)
        uint32_t g1 = g0;
        uint64_t r1 = r0;
        const uint64_t b0 = r0 < nrec ? rec_off[r0] : (nrec ? rec_off[nrec] : 0);
        for (;;) {
            while (r1 < nrec && rec2genome[r1] == g1) r1++;
            g1++;
            const uint64_t b1 = r1 < nrec ? rec_off[r1] : (nrec ? rec_off[nrec] : 0);
            if (g1 >= ngenomes || b1 - b0 >= chunk_bytes) break;
        }
        const uint64_t b1 = r1 < nrec ? rec_off[r1] : (nrec ? rec_off[nrec] : 0);
        const uint64_t base = b0 & ~15ull;                       // keep the 8-byte alignment of the offsets
        const size_t nbytes = (size_t)(b1 - base);
        K1Slot &sl = ws->slot[batch & 1];
        if (sl.used) {                                           // previous batch of this slot must be finished
            if (cudaEventSynchronize(sl.done) != cudaSuccess) { tc_set_error("tc_sketch_minhash: %s", cudaGetErrorString(cudaGetLastError())); rc = 1; break; }
        }
        if (nbytes + 64 > sl.seq_cap) {
            cudaFree(sl.d_seq); sl.d_seq = nullptr; sl.seq_cap = 0;
            if (cudaMalloc(&sl.d_seq, nbytes + 64 + nbytes / 8) != cudaSuccess) { tc_set_error("tc_sketch_minhash: out of device memory (%zu bytes of sequence per batch)", nbytes); rc = 1; break; }
            sl.seq_cap = nbytes + 64 + nbytes / 8;
        }
        // (packed 2-bit transport or plain ASCII copy: k_stage.cuh)
        if (tc_stage_batch(ctx, sl.pack, seqs, base, nbytes, sl.d_seq, cs, st, sl.copied, pack_threads, "tc_sketch_minhash")) { rc = 1; break; }
        k1_plan(ctx, rec_off, rec2genome, r0, r1, g0, g1, k, sketch_size, base, pl);
        rc = k1_enqueue(ctx, sl, st, sl.d_seq, pl, g1 - g0, k, seed, sketch_size + g0, ws->d_out + (size_t)g0 * out_stride,
                        out_stride, ws->d_n + g0);
        if (!rc && cudaEventRecord(sl.done, st) != cudaSuccess) { tc_set_error("tc_sketch_minhash: event record failed"); rc = 1; }
        g0 = g1; r0 = r1; batch++;
    }
    if (!rc) {
        if (cudaMemcpyAsync(out_hashes, ws->d_out, (size_t)ngenomes * out_stride * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(out_n, ws->d_n, (size_t)ngenomes * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
            tc_set_error("tc_sketch_minhash: D2H copy failed");
            rc = 1;
        }
    }
    const cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(st);
    if (!rc && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        tc_set_error("tc_sketch_minhash: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        rc = 1;
    }
    return rc;
}
