// k_stage.cuh -- staging of one sequence batch of the sketch pipelines (K1 MinHash, K2 KSSD) on the device: either the
// ASCII bytes as they are (1 byte per base over PCIe) or the packed 2-bit transport -- the host packs the batch to 2 bits
// per base plus an exception list (pack2bit.cpp), a quarter of the bytes crosses PCIe, and the device rebuilds the exact
// byte stream (north_star: "packed 2-bit k-mer windows"; exact: every byte that is not ACGTacgt travels in the list, a
// lower-case base comes back upper-case, which is what both sketchers hash anyway: MinHash.cpp:472-479, BaseMap of
// SketchInfo.cpp:1100-1125).
#pragma once
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

extern "C" size_t tc_pack2bit_range(const uint8_t *seq, uint64_t i0, uint64_t i1, uint64_t origin, uint8_t *out,
                                    uint32_t *exc_pos, uint8_t *exc_byte, size_t exc_cap);     // pack2bit.cpp

namespace {

// per staging slot: pinned staging of the packed batch and its exception list, and their device copies
struct TcPackBuf {
    uint8_t *h_pack = nullptr, *d_pack = nullptr; size_t pack_cap = 0;
    uint32_t *h_exc_pos = nullptr, *d_exc_pos = nullptr; uint8_t *h_exc_byte = nullptr, *d_exc_byte = nullptr; size_t exc_cap = 0;
    void release()
    {
        cudaFree(d_pack); cudaFree(d_exc_pos); cudaFree(d_exc_byte);
        if (h_pack) cudaFreeHost(h_pack);
        if (h_exc_pos) cudaFreeHost(h_exc_pos);
        if (h_exc_byte) cudaFreeHost(h_exc_byte);
        *this = TcPackBuf();
    }
};

// One thread = 16 bases = one 32-bit word of the packed stream -> one 16-byte store; code -> byte through PRMT look-ups
// in the word 'A','C','T','G'.
__global__ void tc_unpack_kernel(const uint32_t *packed, uint64_t nwords, uint8_t *out)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    const uint32_t p = packed[w];
    const uint32_t table = 0x47544341u;                 // bytes: [0]='A' [1]='C' [2]='T' [3]='G'
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t c = (p >> (8 * q)) & 0xFFu;      // 4 bases
        const uint32_t sel = (c & 3u) | (((c >> 2) & 3u) << 4) | (((c >> 4) & 3u) << 8) | (((c >> 6) & 3u) << 12);
        o[q] = __byte_perm(table, 0u, sel);
    }
    reinterpret_cast<uint4 *>(out)[w] = make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void tc_patch_kernel(const uint32_t *pos, const uint8_t *byte, uint32_t n, uint8_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[pos[i]] = byte[i];
}

// threads the host packer may use: all cores divided by the visible GPUs (one process per GPU shares the host);
// packing pays only when it outruns the PCIe link (~50 GB/s of ASCII; 77 GB/s measured with 16 threads): from 12 threads up.
// TCLUST_SKETCH_PACK=0|1 forces it off / on, TCLUST_PACK_THREADS sets the count.
inline int tc_pack_threads()
{
    const char *e = getenv("TCLUST_SKETCH_PACK");
    if (e && !strcmp(e, "0")) return 0;
    int ndev = 1;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); ndev = 1; }
    int t = (int)std::thread::hardware_concurrency() / ndev;
    if (const char *pt = getenv("TCLUST_PACK_THREADS")) t = atoi(pt);
    if (e && !strcmp(e, "1")) return std::max(t, 1);
    return t >= 12 ? t : 0;
}

// pack seqs[base, base + nbytes) into the pinned staging with `threads` host threads; false: too many non-ACGT bytes
// for the exception list (the batch then goes over as ASCII)
inline bool tc_pack_batch(TcPackBuf &pb, const uint8_t *seqs, uint64_t base, size_t nbytes, int threads, uint32_t *n_exc)
{
    const size_t per = ((nbytes + threads - 1) / threads + 63) & ~(size_t)63;
    const size_t cap_t = pb.exc_cap / threads;
    std::vector<size_t> cnt(threads, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t]() {
            const size_t a = std::min(nbytes, per * t), b = std::min(nbytes, per * (t + 1));
            if (b > a) cnt[t] = tc_pack2bit_range(seqs + base, a, b, 0, pb.h_pack + a / 4, pb.h_exc_pos + cap_t * t, pb.h_exc_byte + cap_t * t, cap_t);
        });
    for (std::thread &x : th) x.join();
    size_t total = 0;
    for (int t = 0; t < threads; t++) {
        if (cnt[t] > cap_t) return false;
        if (t && cnt[t] && total != cap_t * t) {
            memmove(pb.h_exc_pos + total, pb.h_exc_pos + cap_t * t, cnt[t] * 4);
            memmove(pb.h_exc_byte + total, pb.h_exc_byte + cap_t * t, cnt[t]);
        }
        total += cnt[t];
    }
    *n_exc = (uint32_t)total;
    return true;
}

// Bring seqs[base, base + nbytes) (+ 64 zero bytes) into d_seq: copies on `cs`, `copied` recorded behind them, `st` made
// to wait for it; the unpack kernels run on `st`.  The previous use of this slot (its `done` event) must have completed.
// Adds the bytes that cross PCIe to ctx->last_h2d_bytes.  Returns 0 / 1 (error text set, prefixed with `who`).
inline int tc_stage_batch(tc_ctx *ctx, TcPackBuf &pb, const uint8_t *seqs, uint64_t base, size_t nbytes, uint8_t *d_seq,
                          cudaStream_t cs, cudaStream_t st, cudaEvent_t copied, int pack_threads, const char *who)
{
    bool packed = false;
    uint32_t n_exc = 0;
    if (pack_threads > 0 && nbytes >= (1u << 20) && nbytes < (1ull << 32)) {
        const size_t pbytes = ((nbytes + 63) & ~(size_t)63) / 4;
        if (pbytes + 64 > pb.pack_cap) {
            if (pb.h_pack) cudaFreeHost(pb.h_pack);
            cudaFree(pb.d_pack);
            pb.h_pack = nullptr; pb.d_pack = nullptr; pb.pack_cap = 0;
            const size_t cap = pbytes + 64 + pbytes / 8;
            if (cudaMallocHost(&pb.h_pack, cap) != cudaSuccess || cudaMalloc(&pb.d_pack, cap) != cudaSuccess) cudaGetLastError();
            else pb.pack_cap = cap;
        }
        if (!pb.exc_cap) {
            const size_t cap = 1u << 20;
            if (cudaMallocHost(&pb.h_exc_pos, cap * 4) == cudaSuccess && cudaMallocHost(&pb.h_exc_byte, cap) == cudaSuccess &&
                cudaMalloc(&pb.d_exc_pos, cap * 4) == cudaSuccess && cudaMalloc(&pb.d_exc_byte, cap) == cudaSuccess) pb.exc_cap = cap;
            else cudaGetLastError();
        }
        if (pb.pack_cap && pb.exc_cap) packed = tc_pack_batch(pb, seqs, base, nbytes, pack_threads, &n_exc);
        if (packed) {
            const uint64_t nwords = (nbytes + 15) / 16;
            ctx->last_h2d_bytes += pbytes + (uint64_t)n_exc * 5;
            bool ok = cudaMemcpyAsync(pb.d_pack, pb.h_pack, pbytes, cudaMemcpyHostToDevice, cs) == cudaSuccess;
            if (ok && n_exc) ok = cudaMemcpyAsync(pb.d_exc_pos, pb.h_exc_pos, (size_t)n_exc * 4, cudaMemcpyHostToDevice, cs) == cudaSuccess &&
                                  cudaMemcpyAsync(pb.d_exc_byte, pb.h_exc_byte, n_exc, cudaMemcpyHostToDevice, cs) == cudaSuccess;
            ok = ok && cudaEventRecord(copied, cs) == cudaSuccess && cudaStreamWaitEvent(st, copied, 0) == cudaSuccess;
            if (ok) {
                tc_unpack_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint32_t *>(pb.d_pack), nwords, d_seq);
                if (n_exc) tc_patch_kernel<<<(n_exc + 255) / 256, 256, 0, st>>>(pb.d_exc_pos, pb.d_exc_byte, n_exc, d_seq);
                ctx->launches += n_exc ? 2 : 1;
                ok = cudaMemsetAsync(d_seq + nbytes, 0, 64, st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
            }
            if (!ok) { tc_set_error("%s: packed H2D copy failed: %s", who, cudaGetErrorString(cudaGetLastError())); return 1; }
            return 0;
        }
    }
    ctx->last_h2d_bytes += nbytes;
    if ((nbytes && cudaMemcpyAsync(d_seq, seqs + base, nbytes, cudaMemcpyHostToDevice, cs) != cudaSuccess) ||
        cudaMemsetAsync(d_seq + nbytes, 0, 64, cs) != cudaSuccess ||
        cudaEventRecord(copied, cs) != cudaSuccess || cudaStreamWaitEvent(st, copied, 0) != cudaSuccess) {
        tc_set_error("%s: H2D copy failed: %s", who, cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}

}  // namespace
