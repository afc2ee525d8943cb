// k6_dense.cu -- K6: the --dense epilogue of the MST build: per-genome density maps and the ANI histogram over
// the candidate pairs of the dense pass.
//
// Replaces the `if(!no_dense)` block of compute_minhash_mst / compute_kssd_mst (reference src/MST.cpp:1517-1530,
// :546-556) and the cumulative pass (:1702-1712): for every candidate pair (common >= 1, size-ratio filter passed)
// with distance d:  t0 = lower_bound(radius, d), radius[t] = t * (1.0 / span)  ->  dense[t0][i]++, dense[t0][j]++;
// ANI bin (int)((1.0 - d) * 100.0) (clamped to 100) ++;  finally dense[t][v] is made cumulative over t.
//
// Exactness: d is a double produced by glibc log() on the host in the reference.  The device never evaluates log:
// both bin functions are monotone step functions of J = common / denom (the exact IEEE quotient the Boruvka keys
// use), so the host finds, by bisection over double bit patterns with the reference's own expression, the smallest J
// of every step; the kernel bins by comparing J with those thresholds.  An edge whose J lies within 1e-9 (relative)
// of a threshold is not binned on the device but returned for an exact host evaluation (none in practice: achievable
// J are ratios of integers <= 2e4, the thresholds are transcendental).
#include "tc_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

constexpr int K6_MAX_SPAN = 1024;
constexpr uint32_t K6_RECHECK_CAP = 1u << 20;

struct K6Args {
    const tc_pair *edges; unsigned long long n_edges;
    const uint32_t *size; uint32_t n;
    int containment, span;
    const double *dthr;       // span thresholds, non-increasing: smallest J with dist(J) <= t/span
    const double *athr;       // 101 thresholds, non-decreasing: smallest J with ANI(J) >= b
    int *dense;               // span x n
    unsigned long long *ani;  // 101
    tc_pair *recheck; unsigned int *recheck_n;
};

__device__ __forceinline__ bool k6_near(double j, double thr) { return fabs(j - thr) <= 1e-9 * thr; }

__global__ void __launch_bounds__(256) k6_dense_kernel(const K6Args a)
{
    __shared__ double s_dthr[K6_MAX_SPAN];
    __shared__ double s_athr[101];
    __shared__ unsigned long long s_ani[101];
    for (int i = threadIdx.x; i < a.span; i += blockDim.x) s_dthr[i] = a.dthr[i];
    for (int i = threadIdx.x; i < 101; i += blockDim.x) { s_athr[i] = a.athr[i]; s_ani[i] = 0ULL; }
    __syncthreads();
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < a.n_edges;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const tc_pair p = a.edges[e];
        const uint32_t s0 = a.size[p.i], s1 = a.size[p.j];
        const uint32_t denom = a.containment ? min(s0, s1) : (s0 + s1 - p.common);
        const double J = denom == 0 ? 0.0 : __ddiv_rn((double)p.common, (double)denom);
        // density bin: smallest t with J >= dthr[t] (dthr non-increasing); span = none
        int lo = 0, hi = a.span;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (J >= s_dthr[mid]) hi = mid; else lo = mid + 1; }
        const int t0 = lo;
        // ANI bin: largest b with J >= athr[b] (athr[0] = 0)
        int bl = 0, bh = 100;
        while (bl < bh) { const int mid = (bl + bh + 1) >> 1; if (J >= s_athr[mid]) bl = mid; else bh = mid - 1; }
        const int b = bl;
        bool near = false;
        if (t0 < a.span) near |= k6_near(J, s_dthr[t0]);
        if (t0 > 0) near |= k6_near(J, s_dthr[t0 - 1]);
        if (b > 0) near |= k6_near(J, s_athr[b]);
        if (b < 100) near |= k6_near(J, s_athr[b + 1]);
        if (near) {
            const unsigned int idx = atomicAdd(a.recheck_n, 1u);
            if (idx < K6_RECHECK_CAP) a.recheck[idx] = p;
            continue;
        }
        if (t0 < a.span) {
            atomicAdd(&a.dense[(size_t)t0 * a.n + p.i], 1);
            atomicAdd(&a.dense[(size_t)t0 * a.n + p.j], 1);
        }
        atomicAdd(&s_ani[b], 1ULL);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 101; i += blockDim.x)
        if (s_ani[i]) atomicAdd(&a.ani[i], s_ani[i]);
}

// the reference's expressions (MST.cpp:1489-1515) as a function of J
double k6_dist(double J, int containment, int kmer_size)
{
    const double inv_kmer_size = 1.0 / kmer_size;
    if (J == 1.0) return 0.0;
    if (J == 0.0) return 1.0;
    if (containment) return -inv_kmer_size * std::log(J);
    const double ratio = (2.0 * J) / (1.0 + J);
    return -inv_kmer_size * std::log(ratio);
}
int k6_ani_bin(double dist)
{
    const double tmpANI = 1.0 - dist;
    int ANI = (int)(tmpANI * 100.0);
    if (ANI >= 101) ANI = 100;
    return ANI < 0 ? 0 : ANI;
}
// smallest J in (0, 1] with pred(J) true, pred monotone (false ... true) in J; 2.0 if none
template <typename F>
double k6_first_true(F pred)
{
    if (!pred(1.0)) return 2.0;
    uint64_t lo = 1, hi;                      // bit patterns of positive doubles are ordered like the values
    const double one = 1.0;
    memcpy(&hi, &one, 8);
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        double v;
        memcpy(&v, &mid, 8);
        if (pred(v)) hi = mid; else lo = mid + 1;
    }
    double v;
    memcpy(&v, &lo, 8);
    return v;
}

}  // namespace

extern "C" int tc_dense_ani(tc_ctx *ctx, int kmer_size, int containment, int dense_span, int32_t *dense, uint64_t *ani)
{
    TC_CUDA(cudaSetDevice(ctx->device));
    TC_CHECK(ctx->n > 0 && ctx->d_edges, "tc_dense_ani: run tc_allpairs_pass / tc_mst first");
    TC_CHECK(ctx->edges_semb_size == 0, "tc_dense_ani: the candidate edges must come from tc_allpairs_pass (semantic A)");
    TC_CHECK(!ctx->edges_partial, "tc_dense_ani: the last MST was built in bounded edge memory (or by one rank of a multi-GPU group): "
             "the device holds only part of the candidate pairs; run tc_allpairs_pass first");
    TC_CHECK(dense_span >= 1 && dense_span <= K6_MAX_SPAN, "tc_dense_ani: dense_span=%d out of range (1..%d)", dense_span, K6_MAX_SPAN);
    TC_CHECK(containment == 0 || containment == 1, "tc_dense_ani: containment must be 0 or 1");
    const uint32_t n = ctx->n;
    // ---- thresholds in J
    std::vector<double> dthr(dense_span), athr(101, 0.0);
    const double step = 1.0 / dense_span;                              // MST.cpp:1323-1329
    for (int t = 0; t < dense_span; t++) {
        const double radius = step * (double)t;
        dthr[t] = k6_first_true([&](double J) { return k6_dist(J, containment, kmer_size) <= radius; });
    }
    for (int b = 1; b <= 100; b++)
        athr[b] = k6_first_true([&](double J) { return k6_ani_bin(k6_dist(J, containment, kmer_size)) >= b; });
    cudaStream_t st = ctx->stream;
    double *d_thr = nullptr; int *d_dense = nullptr; unsigned long long *d_ani = nullptr; tc_pair *d_re = nullptr;
    unsigned int *d_ren = nullptr;
    auto cleanup = [&]() { cudaFree(d_thr); cudaFree(d_dense); cudaFree(d_ani); cudaFree(d_re); cudaFree(d_ren); };
#define K6_TRY(x) do { if ((x) != cudaSuccess) { tc_set_error("tc_dense_ani: %s failed: %s", #x, cudaGetErrorString(cudaGetLastError())); cleanup(); return 1; } } while (0)
    K6_TRY(cudaMalloc(&d_thr, (size_t)(dense_span + 101) * 8));
    K6_TRY(cudaMalloc(&d_dense, (size_t)dense_span * n * 4));
    K6_TRY(cudaMalloc(&d_ani, 101 * 8));
    K6_TRY(cudaMalloc(&d_re, (size_t)K6_RECHECK_CAP * sizeof(tc_pair)));
    K6_TRY(cudaMalloc(&d_ren, 4));
    K6_TRY(cudaMemcpyAsync(d_thr, dthr.data(), (size_t)dense_span * 8, cudaMemcpyHostToDevice, st));
    K6_TRY(cudaMemcpyAsync(d_thr + dense_span, athr.data(), 101 * 8, cudaMemcpyHostToDevice, st));
    K6_TRY(cudaMemsetAsync(d_dense, 0, (size_t)dense_span * n * 4, st));
    K6_TRY(cudaMemsetAsync(d_ani, 0, 101 * 8, st));
    K6_TRY(cudaMemsetAsync(d_ren, 0, 4, st));
    if (ctx->n_edges) {
        K6Args a{};
        a.edges = ctx->d_edges; a.n_edges = ctx->n_edges; a.size = ctx->d_size; a.n = n;
        a.containment = containment; a.span = dense_span; a.dthr = d_thr; a.athr = d_thr + dense_span;
        a.dense = d_dense; a.ani = d_ani; a.recheck = d_re; a.recheck_n = d_ren;
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((ctx->n_edges + 255) / 256, (uint64_t)ctx->num_sms * 8));
        k6_dense_kernel<<<grid, 256, 0, st>>>(a);
        ctx->launches++;
        K6_TRY(cudaGetLastError());
    }
    unsigned int nre = 0;
    K6_TRY(cudaMemcpyAsync(dense, d_dense, (size_t)dense_span * n * 4, cudaMemcpyDeviceToHost, st));
    K6_TRY(cudaMemcpyAsync(ani, d_ani, 101 * 8, cudaMemcpyDeviceToHost, st));
    K6_TRY(cudaMemcpyAsync(&nre, d_ren, 4, cudaMemcpyDeviceToHost, st));
    K6_TRY(cudaStreamSynchronize(st));
    if (nre > K6_RECHECK_CAP) {
        cleanup();
        tc_set_error("tc_dense_ani: %u candidate pairs sit on a bin boundary, more than the re-check buffer holds", nre);
        return 1;
    }
    if (nre) {                                   // exact host evaluation of the boundary cases
        std::vector<tc_pair> re(nre);
        K6_TRY(cudaMemcpy(re.data(), d_re, (size_t)nre * sizeof(tc_pair), cudaMemcpyDeviceToHost));
        std::vector<double> radius(dense_span);
        for (int t = 0; t < dense_span; t++) radius[t] = step * (double)t;
        for (const tc_pair &p : re) {
            const int s0 = (int)ctx->h_size[p.i], s1 = (int)ctx->h_size[p.j];
            const double d = containment ? tc_aaf_dist((int)p.common, s0, s1, kmer_size) : tc_mash_dist((int)p.common, s0, s1, kmer_size);
            const int t0 = (int)(std::lower_bound(radius.begin(), radius.end(), d) - radius.begin());
            if (t0 < dense_span) { dense[(size_t)t0 * n + p.i]++; dense[(size_t)t0 * n + p.j]++; }
            ani[k6_ani_bin(d)]++;
        }
    }
#undef K6_TRY
    cleanup();
    // start-bucket counts -> cumulative density counts (MST.cpp:1702-1712)
    for (uint32_t v = 0; v < n; v++) {
        int acc = 0;
        for (int t = 0; t < dense_span; t++) { acc += dense[(size_t)t * n + v]; dense[(size_t)t * n + v] = acc; }
    }
    return 0;
}
