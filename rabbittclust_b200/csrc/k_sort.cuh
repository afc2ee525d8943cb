// k_sort.cuh -- block-wide sort / unique / compaction helpers shared by the sketch kernels (K1, K2).
#pragma once
#include <cstdint>

// ------------------------------------------------------------------ block-wide sort / unique
// bitonic sort of buf[0..npad) ascending (npad power of two), all threads of the CTA
static __device__ void cta_bitonic_sort(uint64_t *buf, uint32_t npad)
{
    for (uint32_t size = 2; size <= npad; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1));   // index with bit `stride` clear
                const uint32_t hi = lo + stride;
                const uint64_t a = buf[lo], b = buf[hi];
                const bool up = (lo & size) == 0;
                if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
            }
            __syncthreads();
        }
    }
}

// in-place unique of sorted buf[0..n); returns the number of distinct values (block-uniform).
// scratch: shared uint32_t[40]
static __device__ uint32_t cta_unique_sorted(uint64_t *buf, uint32_t n, uint32_t *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint32_t running = 0;
    uint64_t prev_last = 0;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint64_t v = 0;
        bool keep = false;
        if (i < n) {
            v = buf[i];
            const uint64_t p = (i == base) ? prev_last : buf[i - 1];
            keep = (i == 0) || (v != p);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) scratch[warp] = __popc(bal);
        __syncthreads();                                    // all reads of this round done
        uint32_t before = 0, total = 0;
        for (int w = 0; w < nwarp; w++) { const uint32_t c = scratch[w]; if (w < warp) before += c; total += c; }
        const uint32_t last_i = min(base + blockDim.x, n) - 1;
        if (i == last_i) reinterpret_cast<volatile uint64_t *>(scratch + 34)[0] = v;
        if (keep) buf[running + before + __popc(bal & ((1u << lane) - 1))] = v;
        __syncthreads();
        prev_last = reinterpret_cast<volatile uint64_t *>(scratch + 34)[0];
        running += total;
        __syncthreads();
    }
    return running;
}

// sort + unique + truncate to s; returns new count; *tau_out = s-th smallest when s distinct exist
static __device__ uint32_t cta_compact(uint64_t *buf, uint32_t count, uint32_t cap, uint32_t s, uint32_t *scratch,
                                unsigned long long *tau_out)
{
    uint32_t npad = 2;
    while (npad < count) npad <<= 1;
    if (npad > cap) npad = cap;
    for (uint32_t i = count + threadIdx.x; i < npad; i += blockDim.x) buf[i] = ~0ull;
    __syncthreads();
    cta_bitonic_sort(buf, npad);
    uint32_t m = cta_unique_sorted(buf, count, scratch);
    if (m > s) m = s;
    if (m == s && s > 0) *tau_out = buf[s - 1];
    return m;
}

