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
#include <algorithm>

namespace {

constexpr int K2_THREADS = 256;

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
    int kmer, drlevel, half_outctx_len, dim_end, use64, rev_add_move;
    unsigned long long tupmask, domask, undomask0, undomask1;
    uint64_t *out; unsigned long long out_stride; uint32_t *cnt;
};

__global__ void __launch_bounds__(K2_THREADS) k2_kssd_kernel(const K2Args a)
{
    __shared__ uint32_t s_seg;
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
            if (run >= a.kmer && pos >= e0) {
                const unsigned long long uni = tuple < rvs ? tuple : rvs;
                const uint32_t dim_id = (uint32_t)((uni & a.domask) >> (a.half_outctx_len * 2));
                const int pf = __ldg(a.shuffled_dim + dim_id);
                if (pf < a.dim_end) {
                    unsigned long long dr = (((uni & a.undomask0) |
                                              ((uni & a.undomask1) << (a.kmer * 2 - a.half_outctx_len * 4)))
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

int sketch_kssd_impl(tc_ctx *ctx, const uint8_t *seq_dev, const uint64_t *rec_off, const uint32_t *rec2genome,
                     uint64_t nrec, uint32_t ngenomes, int kmer_size, int drlevel, const int32_t *shuffled_dim,
                     uint64_t *out_dev, uint64_t out_stride, uint32_t *out_n_dev)
{
    TC_CHECK(ngenomes > 0, "tc_sketch_kssd: no genomes");
    TC_CHECK(kmer_size >= 2 && kmer_size <= 32, "tc_sketch_kssd: kmer_size=%d unsupported (2..32)", kmer_size);
    TC_CHECK(out_stride > 0 && out_stride < (1ull << 31), "tc_sketch_kssd: bad out_stride");
    // constants, SketchInfo.cpp:1019-1048
    const int half_k = (kmer_size + 1) / 2;
    const int kmer = half_k * 2;
    const int half_subk = 6 - drlevel >= 2 ? 6 : drlevel + 2;
    TC_CHECK(drlevel >= 0 && half_k > half_subk, "tc_sketch_kssd: half_k=%d must exceed half_subk=%d", half_k, half_subk);
    const int half_outctx_len = half_k - half_subk;
    const int comp_bittl = 64 - 4 * half_k;
    K2Args a{};
    a.kmer = kmer; a.drlevel = drlevel; a.half_outctx_len = half_outctx_len;
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
    if (ctx->kssd_dim_host != shuffled_dim || ctx->kssd_dim_n != dim_n) {
        cudaFree(ctx->d_kssd_dim); ctx->d_kssd_dim = nullptr; ctx->kssd_dim_host = nullptr;
        TC_CUDA(cudaMalloc(&ctx->d_kssd_dim, dim_n * 4));
        TC_CUDA(cudaMemcpyAsync(ctx->d_kssd_dim, shuffled_dim, dim_n * 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->kssd_dim_host = shuffled_dim; ctx->kssd_dim_n = dim_n;
    }
    a.shuffled_dim = ctx->d_kssd_dim;

    // ---- segments of end-positions
    uint64_t total = 0;
    for (uint64_t r = 0; r < nrec; r++) {
        TC_CHECK(rec_off[r + 1] >= rec_off[r], "tc_sketch_kssd: record offsets not monotone");
        TC_CHECK(rec2genome[r] < ngenomes, "tc_sketch_kssd: rec2genome[%llu] out of range", (unsigned long long)r);
        const uint64_t len = rec_off[r + 1] - rec_off[r];
        if (len >= (uint64_t)kmer) total += len - kmer + 1;
    }
    const uint64_t want_ctas = (uint64_t)ctx->num_sms * 32;
    uint64_t segw = std::min<uint64_t>(std::max<uint64_t>(total / std::max<uint64_t>(want_ctas, 1), 1u << 14), 1u << 20);
    segw = (segw + K2_THREADS * 8 - 1) / (K2_THREADS * 8) * (K2_THREADS * 8);
    std::vector<K2Seg> segs;
    for (uint64_t r = 0; r < nrec; r++) {
        const uint64_t len = rec_off[r + 1] - rec_off[r];
        if (len < (uint64_t)kmer) continue;
        const uint64_t first = rec_off[r] + kmer - 1, last = rec_off[r + 1];   // end-positions [first, last)
        for (uint64_t e = first; e < last; e += segw) {
            K2Seg s{};
            s.rec_start = rec_off[r]; s.emit_begin = e; s.emit_end = std::min(e + segw, last); s.genome = rec2genome[r];
            segs.push_back(s);
        }
    }
    const uint32_t nseg = (uint32_t)segs.size();
    cudaStream_t st = ctx->stream;
    K2Seg *d_segs = nullptr; uint32_t *d_cnt = nullptr; unsigned int *d_ctr = nullptr; uint64_t *d_scr = nullptr;
    auto cleanup = [&]() { cudaFree(d_segs); cudaFree(d_cnt); cudaFree(d_ctr); cudaFree(d_scr); };
#define K2_TRY(x) do { if ((x) != cudaSuccess) { tc_set_error("tc_sketch_kssd: %s failed: %s", #x, cudaGetErrorString(cudaGetLastError())); cleanup(); return 1; } } while (0)
    K2_TRY(cudaMalloc(&d_segs, std::max<size_t>(nseg, 1) * sizeof(K2Seg)));
    K2_TRY(cudaMalloc(&d_cnt, (size_t)ngenomes * 4));
    K2_TRY(cudaMalloc(&d_ctr, 8));
    K2_TRY(cudaMemcpyAsync(d_segs, segs.data(), nseg * sizeof(K2Seg), cudaMemcpyHostToDevice, st));
    K2_TRY(cudaMemsetAsync(d_cnt, 0, (size_t)ngenomes * 4, st));
    K2_TRY(cudaMemsetAsync(d_ctr, 0, 8, st));
    a.seq = seq_dev; a.segs = d_segs; a.nseg = nseg; a.seg_counter = d_ctr;
    a.out = out_dev; a.out_stride = out_stride; a.cnt = d_cnt;
    if (nseg) {
        const int grid = (int)std::min<uint64_t>(nseg, (uint64_t)ctx->num_sms * 8);
        k2_kssd_kernel<<<grid, K2_THREADS, 0, st>>>(a);
        ctx->launches++;
        K2_TRY(cudaGetLastError());
    }
    K2SortArgs s{};
    s.out = out_dev; s.out_stride = out_stride; s.cnt = d_cnt; s.out_n = out_n_dev; s.overflow = d_ctr + 1;
    s.cap = pow2_at_least(out_stride);
    size_t smem = (size_t)s.cap * 8;
    s.use_global = smem > 160 * 1024;
    if (s.use_global) {
        K2_TRY(cudaMalloc(&d_scr, (size_t)ngenomes * s.cap * 8));
        s.gscratch = d_scr; smem = 0;
    }
    if (smem > 48 * 1024)
        K2_TRY(cudaFuncSetAttribute(k2_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k2_sort_kernel<<<ngenomes, K2_THREADS, smem, st>>>(s);
    ctx->launches++;
    K2_TRY(cudaGetLastError());
    unsigned int ovf[2] = {0, 0};
    K2_TRY(cudaMemcpyAsync(ovf, d_ctr, 8, cudaMemcpyDeviceToHost, st));
    K2_TRY(cudaStreamSynchronize(st));
#undef K2_TRY
    cleanup();
    TC_CHECK(ovf[1] == 0, "tc_sketch_kssd: a genome produced %u hashes, more than out_stride=%llu", ovf[1],
             (unsigned long long)out_stride);
    return 0;
}

}  // namespace

extern "C" int tc_sketch_kssd_dev(tc_ctx *ctx, const uint8_t *seqs_dev, const uint64_t *rec_off, const uint32_t *rec2genome,
                                  uint64_t nrec, uint32_t ngenomes, int kmer_size, int drlevel,
                                  const int32_t *shuffled_dim, uint64_t *out_hashes_dev, uint64_t out_stride,
                                  uint32_t *out_n_dev)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    return sketch_kssd_impl(ctx, seqs_dev, rec_off, rec2genome, nrec, ngenomes, kmer_size, drlevel, shuffled_dim,
                            out_hashes_dev, out_stride, out_n_dev);
}

extern "C" int tc_sketch_kssd(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off, const uint32_t *rec2genome,
                              uint64_t nrec, uint32_t ngenomes, int kmer_size, int drlevel, const int32_t *shuffled_dim,
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
            tc_set_error("tc_sketch_kssd: out of device memory");
            break;
        }
        if (cudaMemcpyAsync(d_seq, seqs, nbytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemsetAsync(d_seq + nbytes, 0, 64, ctx->stream) != cudaSuccess) {
            tc_set_error("tc_sketch_kssd: H2D copy failed");
            break;
        }
        rc = sketch_kssd_impl(ctx, d_seq, rec_off, rec2genome, nrec, ngenomes, kmer_size, drlevel, shuffled_dim, d_out,
                              out_stride, d_n);
        if (rc) break;
        if (cudaMemcpyAsync(out_hashes, d_out, (uint64_t)ngenomes * out_stride * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(out_n, d_n, (size_t)ngenomes * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            tc_set_error("tc_sketch_kssd: D2H copy failed");
            rc = 1;
        }
    } while (0);
    cudaFree(d_seq); cudaFree(d_out); cudaFree(d_n);
    return rc;
}
