// k2_kssd.cu -- K2: KSSD (--fast) sketch construction, list-mode semantics.
//
// Replaces the hot loop of sketchFileWithKssd (reference src/SketchInfo.cpp:1126-1165, constants
// :1019-1048, per-genome set -> sorted vector :1180-1193).  Per record: 2-bit rolling k-mer `tuple`
// (first base most significant) and its reverse complement `rvs` over kmer = 2*half_k bases; any byte
// outside ACGTacgt restarts the window; canonical = numeric min; the middle 2*half_subk bases select
// a slot of the (host-generated, glibc-rand) shuffled_dim table; the k-mer is kept iff the slot value
// pf < 16^(half_subk-drlevel); hash = outer context bits shifted down by 4*drlevel, or-ed with pf.
// The genome's sketch is the ascending set of distinct hashes.
//
// Design (DESIGN.md "K2"): rolling state lives in registers; every thread walks its own contiguous run
// of bases (8 bytes per load, K1's scheme) and primes itself with the kmer-1 bases in front of its
// run, so there is no cross-thread state.  Only ~1/16^drlevel of the windows survive the filter, so
// survivors are appended to the genome's output row with one global atomic each; a second kernel
// sorts/deduplicates every row in shared memory (block bitonic sort).  The shuffled_dim table (64 MB
// at half_subk = 6) is read through L2 (it fits the 126 MB L2).
#include "tc_common.cuh"
#include "k_sort.cuh"
#include "k_stage.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int K2_THREADS = 256;
// presence filter of the dimension test.  Only 16^(half_subk-drlevel) of the 16^half_subk slots pass
// (4096 of 16.7 M at the defaults); testing every window against the 64 MB table makes the kernel L2-sector bound
// (one random 32 B sector per base).  A 2^18-bit bitmap in shared memory (bit = slot id mod 2^18) answers
// "cannot pass" for 98.4 % of the windows; only the rest read the table.
constexpr int K2_FILTER_BITS = 1 << 18;
constexpr int K2_FILTER_WORDS = K2_FILTER_BITS / 32;

struct K2Seg {
    unsigned long long rec_start;   // first byte of the record
    unsigned long long emit_begin;  // first end-position (absolute byte offset) this segment owns
    unsigned long long emit_end;    // one past the last end-position
    uint32_t genome;
    uint32_t pad;
};

struct K2Args {
    const uint8_t *seq;
    const K2Seg *segs;
    uint32_t nseg;
    unsigned int *seg_counter;
    const int32_t *shuffled_dim;
    const uint32_t *dim_filter;     // K2_FILTER_BITS-bit presence bitmap of the slots that pass (see k2_setup)
    int kmer, drlevel, half_outctx_len, dim_end, use64, rev_add_move;
    int odd;          // 1: per-record mode of the single-file consumer with an odd k-mer size (see tc_sketch_kssd_seq)
    int dr_shift;     // left shift of the low outer-context bits: 2*kmerSize - 4*half_outctx_len
    unsigned long long tupmask, domask, undomask0, undomask1;
    uint64_t *out; unsigned long long out_stride; uint32_t *cnt;
};

// ODD: the per-record mode with an odd k-mer size (a.odd); a template parameter so that the list-mode instantiation -- an
// ALU-bound loop -- carries none of its instructions
template <bool ODD>
__global__ void __launch_bounds__(K2_THREADS) k2_kssd_kernel(const K2Args a)
{
    __shared__ uint32_t s_seg;
    __shared__ uint32_t s_filter[K2_FILTER_WORDS];
    for (int i = threadIdx.x; i < K2_FILTER_WORDS; i += K2_THREADS) s_filter[i] = __ldg(a.dim_filter + i);
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) s_seg = atomicAdd(a.seg_counter, 1u);
        __syncthreads();
        const uint32_t seg_id = s_seg;
        __syncthreads();
        if (seg_id >= a.nseg) break;
        const K2Seg seg = a.segs[seg_id];
        const unsigned long long span = seg.emit_end - seg.emit_begin;
        unsigned long long R = (span + K2_THREADS - 1) / K2_THREADS;
        R = (R + 7) & ~7ULL;
        const unsigned long long e0 = seg.emit_begin + (unsigned long long)threadIdx.x * R;
        if (e0 >= seg.emit_end) continue;
        const unsigned long long e1 = min(e0 + R, seg.emit_end);
        // prime with the kmer-1 bases in front of the first owned end-position (never before the record)
        unsigned long long pos = e0 - seg.rec_start >= (unsigned long long)(a.kmer - 1) ? e0 - (a.kmer - 1) : seg.rec_start;
        const uint64_t *wp = reinterpret_cast<const uint64_t *>(a.seq + (pos & ~7ULL));
        uint64_t word = __ldg(wp) >> (8 * (pos & 7ULL));
        int left = 8 - (int)(pos & 7ULL);
        unsigned long long tuple = 0, rvs = 0;
        int run = 0;                                        // consecutive valid bases so far (reference: base-1)
        uint64_t *row = a.out + (size_t)seg.genome * a.out_stride;
        while (pos < e1) {
            if (left == 0) { word = __ldg(++wp); left = 8; }
            const uint32_t c = (uint32_t)word & 0xDFu;      // fold case
            word >>= 8; left--;
            const uint32_t x = c - 0x41u;                   // 'A'
            const bool valid = x < 20u && ((0x80045u >> x) & 1u);   // A, C, G, T
            if (!valid) { run = 0; tuple = 0; rvs = 0; pos++; continue; }
            uint32_t bn = (c >> 1) & 3u;                    // A0 C1 G3 T2
            bn ^= bn >> 1;                                  // A0 C1 G2 T3
            tuple = ((tuple << 2) | bn) & a.tupmask;
            rvs = (rvs >> 2) + ((unsigned long long)(bn ^ 3u) << a.rev_add_move);
            run++;
            unsigned long long t_use = tuple, r_use = rvs;
            bool use = run >= a.kmer && pos >= e0;
            if (ODD && run == a.kmer - 1 && pos >= e0) {
                // consumer_fasta_task_with_kssd with kmerSize = a.kmer - 1 (SketchInfo.cpp:358-372): the window is used
                // after kmerSize valid bases while the tuple holds a.kmer, and a non-ACGT byte does not clear the tuple:
                // the oldest slot still holds the last valid base before the gap (nothing at the start of the record).
                // This thread saw the byte in front of the run (priming starts a.kmer-1 bytes before e0), so the run
                // really began at a reset; walk back over the gap for that base -- once per reset.
                unsigned long long q = pos - (unsigned long long)(a.kmer - 2);         // first base of the run
                int stale = -1;
                while (q > seg.rec_start && stale < 0) {
                    q--;
                    const uint32_t cc = (uint32_t)__ldg(a.seq + q) & 0xDFu;
                    const uint32_t xx = cc - 0x41u;
                    if (xx < 20u && ((0x80045u >> xx) & 1u)) { uint32_t b = (cc >> 1) & 3u; b ^= b >> 1; stale = (int)b; }
                }
                if (stale >= 0) {
                    t_use = tuple | ((unsigned long long)stale << (2 * (a.kmer - 1)));
                    r_use = rvs | (unsigned long long)(stale ^ 3);
                }
                use = true;
            }
            if (use) {
                const unsigned long long uni = t_use < r_use ? t_use : r_use;
                const uint32_t dim_id = (uint32_t)((uni & a.domask) >> (a.half_outctx_len * 2));
                const uint32_t fb = dim_id & (K2_FILTER_BITS - 1);
                const bool maybe = (s_filter[fb >> 5] >> (fb & 31u)) & 1u;
                const int pf = maybe ? __ldg(a.shuffled_dim + dim_id) : a.dim_end;
                if (pf < a.dim_end) {
                    unsigned long long dr = (((uni & a.undomask0) |
                                              ((uni & a.undomask1) << a.dr_shift))
                                             >> (a.drlevel * 4)) | (unsigned long long)pf;
                    if (!a.use64) dr &= 0xffffffffULL;
                    const uint32_t idx = atomicAdd(&a.cnt[seg.genome], 1u);
                    if (idx < a.out_stride) row[idx] = dr;
                }
            }
            pos++;
        }
    }
}

// per genome: sort + unique the appended hashes in place
struct K2SortArgs {
    uint64_t *out; unsigned long long out_stride; const uint32_t *cnt; uint32_t *out_n;
    uint32_t cap; int use_global; uint64_t *gscratch; unsigned int *overflow;
};

__global__ void __launch_bounds__(K2_THREADS) k2_sort_kernel(const K2SortArgs a)
{
    extern __shared__ __align__(16) uint8_t k2_smem[];
    __shared__ __align__(8) uint32_t s_scratch[40];
    const uint32_t g = blockIdx.x;
    uint64_t *row = a.out + (size_t)g * a.out_stride;
    uint32_t n = a.cnt[g];
    if (n > a.out_stride) {
        if (threadIdx.x == 0) atomicMax(a.overflow, n);
        n = (uint32_t)a.out_stride;
    }
    uint64_t *buf = a.use_global ? a.gscratch + (size_t)g * a.cap : reinterpret_cast<uint64_t *>(k2_smem);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) buf[i] = row[i];
    __syncthreads();
    unsigned long long dummy = 0;
    const uint32_t m = n ? cta_compact(buf, n, a.cap, 0xffffffffu, s_scratch, &dummy) : 0u;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) row[i] = buf[i];
    if (threadIdx.x == 0) a.out_n[g] = m;
}

uint32_t pow2_at_least(uint64_t x) { uint32_t p = 2; while (p < x) p <<= 1; return p; }

// ------------------------------------------------------------------------------ host driver
// Same scheme as K1 (k1_sketch.cu): a batch of whole genomes is planned on the host, its metadata goes through a
// pinned staging buffer, two kernels are enqueued without synchronising; the host-buffer entry point alternates
// batches over two slots so that the H2D copy of batch c+1 overlaps the kernels of batch c.
struct K2Slot {
    uint8_t *d_seq = nullptr;  size_t seq_cap = 0;
    uint8_t *h_meta = nullptr; size_t h_meta_cap = 0;   // pinned: segments
    uint8_t *d_meta = nullptr; size_t d_meta_cap = 0;   // segments, per-genome counters, segment counter, scratch
    cudaEvent_t copied = nullptr, done = nullptr;
    bool used = false;
    TcPackBuf pack;                                      // packed 2-bit transport (k_stage.cuh)
};
struct K2Workspace {
    K2Slot slot[2];
    cudaStream_t copy_stream = nullptr;
    uint64_t *d_out = nullptr; size_t out_cap = 0;
    uint32_t *d_n = nullptr;   size_t n_cap = 0;
    unsigned int *d_overflow = nullptr;                  // max pre-deduplication row length seen in a call
};

K2Workspace *k2_ws(tc_ctx *ctx)
{
    if (!ctx->k2_ws) {
        K2Workspace *w = new K2Workspace();
        if (cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc(&w->d_overflow, 4) != cudaSuccess) { delete w; return nullptr; }
        for (K2Slot &sl : w->slot) {
            cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
        }
        ctx->k2_ws = w;
    }
    return static_cast<K2Workspace *>(ctx->k2_ws);
}

inline size_t k2_al16(size_t x) { return (x + 15) & ~(size_t)15; }

// constants of the scheme (SketchInfo.cpp:1019-1048) + the resident shuffled_dim table
int k2_setup(tc_ctx *ctx, int kmer_size, int drlevel, const int32_t *shuffled_dim, uint64_t out_stride, K2Args &a, bool seq_mode = false)
{
    TC_CHECK(kmer_size >= 2 && kmer_size <= 32, "tc_sketch_kssd: kmer_size=%d unsupported (2..32)", kmer_size);
    TC_CHECK(out_stride > 0 && out_stride < (1ull << 31), "tc_sketch_kssd: bad out_stride");
    const int half_k = (kmer_size + 1) / 2;
    const int kmer = half_k * 2;
    const int half_subk = 6 - drlevel >= 2 ? 6 : drlevel + 2;
    TC_CHECK(drlevel >= 0 && half_k > half_subk, "tc_sketch_kssd: half_k=%d must exceed half_subk=%d", half_k, half_subk);
    const int half_outctx_len = half_k - half_subk;
    const int comp_bittl = 64 - 4 * half_k;
    a.kmer = kmer; a.drlevel = drlevel; a.half_outctx_len = half_outctx_len;
    // list mode rounds kmerSize up to 2*half_k (SketchInfo.cpp:1019-1020); the single-file consumer does not (:296, :386)
    a.odd = seq_mode && (kmer_size & 1);
    a.dr_shift = (a.odd ? kmer_size : kmer) * 2 - half_outctx_len * 4;
    a.dim_end = 1 << (4 * (half_subk - drlevel));
    a.use64 = half_k - drlevel > 8;
    a.rev_add_move = 4 * half_k - 2;
    a.tupmask = UINT64_MAX >> comp_bittl;
    a.domask = (a.tupmask >> (4 * half_outctx_len)) << (2 * half_outctx_len);
    const unsigned long long undomask = (a.tupmask ^ a.domask) & a.tupmask;
    a.undomask1 = undomask & (a.tupmask >> ((half_k + half_subk) * 2));
    a.undomask0 = undomask ^ a.undomask1;
    // shuffled_dim table: resident per context (the reference generates it once per process)
    const size_t dim_n = (size_t)1 << (4 * half_subk);
    if (ctx->kssd_dim_host != shuffled_dim || ctx->kssd_dim_n != dim_n || ctx->kssd_dim_end != a.dim_end) {
        cudaFree(ctx->d_kssd_dim); ctx->d_kssd_dim = nullptr; ctx->kssd_dim_host = nullptr;
        cudaFree(ctx->d_kssd_filter); ctx->d_kssd_filter = nullptr;
        TC_CUDA(cudaMalloc(&ctx->d_kssd_dim, dim_n * 4));
        TC_CUDA(cudaMalloc(&ctx->d_kssd_filter, K2_FILTER_WORDS * 4));
        TC_CUDA(cudaMemcpy(ctx->d_kssd_dim, shuffled_dim, dim_n * 4, cudaMemcpyHostToDevice));
        std::vector<uint32_t> filt(K2_FILTER_WORDS, 0u);
        for (size_t t = 0; t < dim_n; t++)
            if (shuffled_dim[t] < a.dim_end) { const uint32_t fb = (uint32_t)t & (K2_FILTER_BITS - 1); filt[fb >> 5] |= 1u << (fb & 31u); }
        TC_CUDA(cudaMemcpy(ctx->d_kssd_filter, filt.data(), K2_FILTER_WORDS * 4, cudaMemcpyHostToDevice));
        ctx->kssd_dim_host = shuffled_dim; ctx->kssd_dim_n = dim_n; ctx->kssd_dim_end = a.dim_end;
    }
    a.shuffled_dim = ctx->d_kssd_dim;
    a.dim_filter = ctx->d_kssd_filter;
    return 0;
}

int k2_validate(const uint64_t *rec_off, const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes)
{
    TC_CHECK(ngenomes > 0, "tc_sketch_kssd: no genomes");
    for (uint64_t r = 0; r < nrec; r++) {
        TC_CHECK(rec_off[r + 1] >= rec_off[r], "tc_sketch_kssd: record offsets not monotone");
        TC_CHECK(r == 0 || rec2genome[r] >= rec2genome[r - 1], "tc_sketch_kssd: rec2genome must be non-decreasing");
        TC_CHECK(rec2genome[r] < ngenomes, "tc_sketch_kssd: rec2genome[%llu] out of range", (unsigned long long)r);
    }
    return 0;
}

// enqueue one batch: records [r0, r1) = genomes [g0, g0+ngl); byte x of the input lives at seq_dev[x - base_off];
// out_dev / out_n_dev point at genome g0
int k2_enqueue(tc_ctx *ctx, K2Workspace *ws, K2Slot &sl, cudaStream_t st, K2Args a, const uint8_t *seq_dev,
               const uint64_t *rec_off, const uint32_t *rec2genome, uint64_t r0, uint64_t r1, uint32_t g0, uint32_t ngl,
               uint64_t base_off, uint64_t *out_dev, uint64_t out_stride, uint32_t *out_n_dev)
{
    uint64_t total = 0;
    const int kmin = a.kmer - a.odd;                     // shortest run of valid bases that yields a window
    for (uint64_t r = r0; r < r1; r++) {
        const uint64_t len = rec_off[r + 1] - rec_off[r];
        if (len >= (uint64_t)kmin) total += len - kmin + 1;
    }
    const uint64_t want_ctas = (uint64_t)ctx->num_sms * 32;
    uint64_t segw = std::min<uint64_t>(std::max<uint64_t>(total / std::max<uint64_t>(want_ctas, 1), 1u << 14), 1u << 20);
    segw = (segw + K2_THREADS * 8 - 1) / (K2_THREADS * 8) * (K2_THREADS * 8);
    std::vector<K2Seg> segs;
    for (uint64_t r = r0; r < r1; r++) {
        const uint64_t len = rec_off[r + 1] - rec_off[r];
        if (len < (uint64_t)kmin) continue;
        const uint64_t start = rec_off[r] - base_off;
        const uint64_t first = start + kmin - 1, last = start + len;        // end-positions [first, last)
        for (uint64_t e = first; e < last; e += segw) {
            K2Seg sg{};
            sg.rec_start = start; sg.emit_begin = e; sg.emit_end = std::min(e + segw, last); sg.genome = rec2genome[r] - g0;
            segs.push_back(sg);
        }
    }
    const uint32_t nseg = (uint32_t)segs.size();
    K2SortArgs s{};
    s.cap = pow2_at_least(out_stride);
    size_t smem = (size_t)s.cap * 8;
    s.use_global = smem > 160 * 1024;
    const size_t host_bytes = k2_al16(std::max<size_t>(nseg, 1) * sizeof(K2Seg));
    const size_t o_cnt = host_bytes;
    const size_t o_ctr = k2_al16(o_cnt + (size_t)ngl * 4);
    size_t total_bytes = k2_al16(o_ctr + 16);
    size_t o_scr = 0;
    if (s.use_global) { o_scr = total_bytes; total_bytes = k2_al16(total_bytes + (size_t)ngl * s.cap * 8); smem = 0; }
    if (host_bytes > sl.h_meta_cap) {
        if (sl.h_meta) cudaFreeHost(sl.h_meta);
        sl.h_meta = nullptr; sl.h_meta_cap = 0;
        TC_CUDA(cudaMallocHost(&sl.h_meta, host_bytes * 2));
        sl.h_meta_cap = host_bytes * 2;
    }
    if (total_bytes > sl.d_meta_cap) {
        if (sl.d_meta) cudaFree(sl.d_meta);
        sl.d_meta = nullptr; sl.d_meta_cap = 0;
        TC_CUDA(cudaMalloc(&sl.d_meta, total_bytes + total_bytes / 4));
        sl.d_meta_cap = total_bytes + total_bytes / 4;
    }
    if (nseg) memcpy(sl.h_meta, segs.data(), (size_t)nseg * sizeof(K2Seg));
    TC_CUDA(cudaMemcpyAsync(sl.d_meta, sl.h_meta, host_bytes, cudaMemcpyHostToDevice, st));
    TC_CUDA(cudaMemsetAsync(sl.d_meta + o_cnt, 0, o_ctr + 16 - o_cnt, st));
    uint32_t *d_cnt = reinterpret_cast<uint32_t *>(sl.d_meta + o_cnt);
    a.seq = seq_dev; a.segs = reinterpret_cast<const K2Seg *>(sl.d_meta); a.nseg = nseg;
    a.seg_counter = reinterpret_cast<unsigned int *>(sl.d_meta + o_ctr);
    a.out = out_dev; a.out_stride = out_stride; a.cnt = d_cnt;
    if (nseg) {
        const int grid = (int)std::min<uint64_t>(nseg, (uint64_t)ctx->num_sms * 6);   // 6 CTAs x 32 KB filter per SM
        if (a.odd) k2_kssd_kernel<true><<<grid, K2_THREADS, 0, st>>>(a);
        else k2_kssd_kernel<false><<<grid, K2_THREADS, 0, st>>>(a);
        ctx->launches++;
        TC_CUDA(cudaGetLastError());
    }
    s.out = out_dev; s.out_stride = out_stride; s.cnt = d_cnt; s.out_n = out_n_dev; s.overflow = ws->d_overflow;
    s.gscratch = s.use_global ? reinterpret_cast<uint64_t *>(sl.d_meta + o_scr) : nullptr;
    if (smem > 48 * 1024)
        TC_CUDA(cudaFuncSetAttribute(k2_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k2_sort_kernel<<<ngl, K2_THREADS, smem, st>>>(s);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    sl.used = true;
    return 0;
}

static thread_local uint32_t g_kssd_needed = 0;

int k2_check_overflow(K2Workspace *ws, cudaStream_t st, uint64_t out_stride)
{
    unsigned int ovf = 0;
    TC_CUDA(cudaMemcpyAsync(&ovf, ws->d_overflow, 4, cudaMemcpyDeviceToHost, st));
    TC_CUDA(cudaStreamSynchronize(st));
    if (ovf != 0) {      // distinct return code; the needed row length is in the message AND in tc_kssd_needed_stride()
        tc_set_error("tc_sketch_kssd: a genome produced %u hashes, more than out_stride=%llu", ovf, (unsigned long long)out_stride);
        g_kssd_needed = ovf;
        return TC_ERR_KSSD_STRIDE;
    }
    return 0;
}

}  // namespace

void tc_k2_free(tc_ctx *ctx)
{
    K2Workspace *w = static_cast<K2Workspace *>(ctx->k2_ws);
    if (!w) return;
    for (K2Slot &sl : w->slot) {
        cudaFree(sl.d_seq); cudaFree(sl.d_meta);
        sl.pack.release();
        if (sl.h_meta) cudaFreeHost(sl.h_meta);
        if (sl.copied) cudaEventDestroy(sl.copied);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    cudaFree(w->d_out); cudaFree(w->d_n); cudaFree(w->d_overflow);
    if (w->copy_stream) cudaStreamDestroy(w->copy_stream);
    delete w;
    ctx->k2_ws = nullptr;
}

// generate_shuffle_dim (reference src/SketchInfo.cpp:60-102): identity permutation of 16^half_subk slots,
// Fisher-Yates driven by glibc srand(23)/rand(), then again by srand(348842630).  The glibc rand() stream is part
// of the reference's definition of a KSSD sketch, so this stays on the host.
static void kssd_shuffle(int32_t *arr, int length, unsigned seed)
{
    srand(seed);
    for (int i = length - 1; i > 0; i--) {
        const int j = rand() % (i + 1);
        const int32_t tmp = arr[i]; arr[i] = arr[j]; arr[j] = tmp;
    }
}
extern "C" int tc_kssd_shuffled_dim(int half_subk, int32_t *out)
{
    TC_CHECK(half_subk >= 1 && half_subk <= 7, "tc_kssd_shuffled_dim: half_subk=%d out of range (1..7)", half_subk);
    const int n = 1 << (4 * half_subk);
    for (int i = 0; i < n; i++) out[i] = i;
    kssd_shuffle(out, n, 23);
    kssd_shuffle(out, n, 348842630u);
    return 0;
}

extern "C" int tc_sketch_kssd_dev(tc_ctx *ctx, const uint8_t *seqs_dev, const uint64_t *rec_off, const uint32_t *rec2genome,
                                  uint64_t nrec, uint32_t ngenomes, int kmer_size, int drlevel,
                                  const int32_t *shuffled_dim, uint64_t *out_hashes_dev, uint64_t out_stride,
                                  uint32_t *out_n_dev)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    if (k2_validate(rec_off, rec2genome, nrec, ngenomes)) return 1;
    K2Workspace *ws = k2_ws(ctx);
    TC_CHECK(ws, "tc_sketch_kssd: cannot create the workspace");
    K2Args a{};
    if (k2_setup(ctx, kmer_size, drlevel, shuffled_dim, out_stride, a)) return 1;
    TC_CUDA(cudaMemsetAsync(ws->d_overflow, 0, 4, ctx->stream));
    if (k2_enqueue(ctx, ws, ws->slot[0], ctx->stream, a, seqs_dev, rec_off, rec2genome, 0, nrec, 0, ngenomes, 0,
                   out_hashes_dev, out_stride, out_n_dev)) return 1;
    return k2_check_overflow(ws, ctx->stream, out_stride);
}

static int k2_sketch_host(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off, const uint32_t *rec2genome,
                          uint64_t nrec, uint32_t ngenomes, int kmer_size, int drlevel, const int32_t *shuffled_dim,
                          uint64_t *out_hashes, uint64_t out_stride, uint32_t *out_n, bool seq_mode)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    if (k2_validate(rec_off, rec2genome, nrec, ngenomes)) return 1;
    K2Workspace *ws = k2_ws(ctx);
    TC_CHECK(ws, "tc_sketch_kssd: cannot create the workspace");
    K2Args a{};
    if (k2_setup(ctx, kmer_size, drlevel, shuffled_dim, out_stride, a, seq_mode)) return 1;
    cudaStream_t st = ctx->stream, cs = ws->copy_stream;
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
    TC_CUDA(cudaMemsetAsync(ws->d_overflow, 0, 4, st));
    const int pack_threads = tc_pack_threads();
    ctx->last_h2d_bytes = 0;
    uint64_t chunk_bytes = 128ull << 20;
    if (const char *e = getenv("TCLUST_SKETCH_CHUNK_MB")) chunk_bytes = std::max<uint64_t>(strtoull(e, nullptr, 10), 1) << 20;
    const uint64_t end_off = nrec ? rec_off[nrec] : 0;
    uint64_t r0 = 0;
    uint32_t g0 = 0, batch = 0;
    int rc = 0;
    while (g0 < ngenomes && !rc) {
        uint32_t g1 = g0;
        uint64_t r1 = r0;
        const uint64_t b0 = r0 < nrec ? rec_off[r0] : end_off;
        for (;;) {
            while (r1 < nrec && rec2genome[r1] == g1) r1++;
            g1++;
            const uint64_t b1 = r1 < nrec ? rec_off[r1] : end_off;
            if (g1 >= ngenomes || b1 - b0 >= chunk_bytes) break;
        }
        const uint64_t b1 = r1 < nrec ? rec_off[r1] : end_off;
        const uint64_t base = b0 & ~15ull;
        const size_t nbytes = (size_t)(b1 - base);
        K2Slot &sl = ws->slot[batch & 1];
        if (sl.used && cudaEventSynchronize(sl.done) != cudaSuccess) { tc_set_error("tc_sketch_kssd: %s", cudaGetErrorString(cudaGetLastError())); rc = 1; break; }
        if (nbytes + 64 > sl.seq_cap) {
            cudaFree(sl.d_seq); sl.d_seq = nullptr; sl.seq_cap = 0;
            if (cudaMalloc(&sl.d_seq, nbytes + 64 + nbytes / 8) != cudaSuccess) { tc_set_error("tc_sketch_kssd: out of device memory (%zu bytes of sequence per batch)", nbytes); rc = 1; break; }
            sl.seq_cap = nbytes + 64 + nbytes / 8;
        }
        if (tc_stage_batch(ctx, sl.pack, seqs, base, nbytes, sl.d_seq, cs, st, sl.copied, pack_threads, "tc_sketch_kssd")) { rc = 1; break; }
        rc = k2_enqueue(ctx, ws, sl, st, a, sl.d_seq, rec_off, rec2genome, r0, r1, g0, g1 - g0, base,
                        ws->d_out + (size_t)g0 * out_stride, out_stride, ws->d_n + g0);
        if (!rc && cudaEventRecord(sl.done, st) != cudaSuccess) { tc_set_error("tc_sketch_kssd: event record failed"); rc = 1; }
        g0 = g1; r0 = r1; batch++;
    }
    if (!rc) {
        if (cudaMemcpyAsync(out_hashes, ws->d_out, (size_t)ngenomes * out_stride * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(out_n, ws->d_n, (size_t)ngenomes * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
            tc_set_error("tc_sketch_kssd: D2H copy failed");
            rc = 1;
        }
    }
    const cudaError_t e1 = cudaStreamSynchronize(cs);
    if (!rc) rc = k2_check_overflow(ws, st, out_stride);
    else cudaStreamSynchronize(st);
    if (!rc && e1 != cudaSuccess) { tc_set_error("tc_sketch_kssd: %s", cudaGetErrorString(e1)); rc = 1; }
    return rc;
}

extern "C" int tc_sketch_kssd(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off, const uint32_t *rec2genome,
                              uint64_t nrec, uint32_t ngenomes, int kmer_size, int drlevel, const int32_t *shuffled_dim,
                              uint64_t *out_hashes, uint64_t out_stride, uint32_t *out_n)
{
    return k2_sketch_host(ctx, seqs, rec_off, rec2genome, nrec, ngenomes, kmer_size, drlevel, shuffled_dim, out_hashes, out_stride,
                          out_n, false);
}

// one sketch per record of a multi-FASTA: consumer_fasta_task_with_kssd (SketchInfo.cpp:274-433).  rec2genome[r] == r.
extern "C" int tc_sketch_kssd_seq(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off, const uint32_t *rec2genome,
                                  uint64_t nrec, uint32_t ngenomes, int kmer_size, int drlevel, const int32_t *shuffled_dim,
                                  uint64_t *out_hashes, uint64_t out_stride, uint32_t *out_n)
{
    TC_CHECK(nrec == ngenomes, "tc_sketch_kssd_seq: one record per sketch (nrec=%llu, ngenomes=%u)", (unsigned long long)nrec, ngenomes);
    return k2_sketch_host(ctx, seqs, rec_off, rec2genome, nrec, ngenomes, kmer_size, drlevel, shuffled_dim, out_hashes, out_stride,
                          out_n, true);
}

extern "C" uint32_t tc_kssd_needed_stride(void) { return g_kssd_needed; }
