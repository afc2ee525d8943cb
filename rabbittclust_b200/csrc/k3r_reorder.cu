// k3r_reorder.cu -- reordered shadow copy of the resident sketch set for the dense all-pairs pass
//
// The block join (k3j_block_kernel, k3h_join.cu) is fast when the 32 sketches of a block are related: their union
// list is a fraction of their elements and the row table fits one or two partitions.  Whether neighbours in the
// reference's input order (the genome list file, MST.cpp:1399-1437 walks it as given) are related is an accident of
// how the list was written.  When they are not, the upload builds a second copy of the rank arrays with the sketches
// sorted by a locality key -- the smallest rank the sketch shares with any other sketch, i.e. its MinHash over the
// hashes that can contribute to an intersection at all -- so that related sketches sit in the same blocks.  The dense
// pass then runs on that copy and the (i, j) of every candidate edge are mapped back through the permutation.  The
// edge SET is the same as the one of the unpermuted pass (each unordered pair is visited once either way, and
// common(i, j) does not depend on the order); nothing downstream depends on the order of the edge list, which the
// atomic appends leave unspecified anyway.
#include "tc_common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

// locality key of sketch i: the first of its leading ranks that occurs in another sketch too (bitmap `multi` of the
// rank transform)
__global__ void k3r_key_kernel(const uint32_t *ranks, const unsigned long long *off, const uint32_t *multi, uint32_t n,
                               uint32_t *key, uint32_t *idx)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    idx[i] = i;
    const unsigned long long o = off[i], len = off[i + 1] - o;
    uint32_t k = 0xFFFFFFFFu;
    if (len) {
        k = ranks[o];
        const unsigned long long lim = len < 256 ? len : 256;
        for (unsigned long long e = 0; e < lim; e++) {
            const uint32_t r = ranks[o + e];
            if ((multi[r >> 5] >> (r & 31)) & 1u) { k = r; break; }
        }
    }
    key[i] = k;
}

// one CTA per sketch of the copy
__global__ void k3r_gather_kernel(const uint32_t *ranks, const unsigned long long *off, const uint32_t *perm,
                                  const unsigned long long *new_off, uint32_t n, uint32_t *out, uint32_t *new_size)
{
    const uint32_t i = blockIdx.x;
    if (i >= n) return;
    const unsigned long long src = off[perm[i]], dst = new_off[i], len = new_off[i + 1] - dst;
    for (unsigned long long e = threadIdx.x; e < len; e += blockDim.x) out[dst + e] = ranks[src + e];
    if (threadIdx.x == 0) new_size[i] = (uint32_t)len;
}

__global__ void k3r_remap_kernel(tc_pair *edges, const unsigned long long *counters, unsigned long long cap, const uint32_t *perm,
                                 unsigned long long first)
{
    unsigned long long m = counters[0];
    if (m > cap) m = cap;
    for (unsigned long long e = first + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < m;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        tc_pair p = edges[e];
        const uint32_t a = perm[p.i], b = perm[p.j];
        p.i = a > b ? a : b; p.j = a > b ? b : a;               // the pass reports j < i
        edges[e] = p;
    }
}

template <typename T>
int grow(T **p, uint64_t *cap, uint64_t need)
{
    if (*p && need <= *cap) return 0;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    TC_CUDA(cudaMalloc(p, std::max<uint64_t>(need, 1) * sizeof(T)));
    *cap = need;
    return 0;
}

}  // namespace

void tc_shadow_free(tc_ctx *ctx)
{
    tc_ctx *s = ctx->shadow;
    if (!s) return;
    cudaFree(s->d_ranks); cudaFree(s->d_off); cudaFree(s->d_size); cudaFree(s->d_perm);
    cudaFree(s->d_bu_keys); cudaFree(s->d_bu_mask); cudaFree(s->d_bu_count);
    cudaFree(s->d_kh_pivots); cudaFree(s->d_kh_pivoff); cudaFree(s->d_kh_units); cudaFree(s->d_kj_units); cudaFree(s->d_scratch);
    cudaFree(s->d_tile_flags); cudaFree(s->d_tile_rowcount); cudaFree(s->d_ks_tiles); cudaFree(s->d_ks_count);
    delete s;                      // stream, events, edge buffer, counters and column state belong to the owner
    ctx->shadow = nullptr; ctx->use_shadow = false;
}

// share of the elements the union lists keep (1.0: nothing shared inside the blocks); blocks without a list count in full
static double list_ratio(const tc_ctx *c)
{
    uint64_t keys = 0;
    const uint32_t nb = (c->n + 31) / 32;
    for (uint32_t b = 0; b < nb && b < c->bu_count.size(); b++) {
        const uint64_t elems = c->h_off[std::min<uint32_t>(c->n, b * 32 + 32)] - c->h_off[b * 32];
        keys += c->bu_count[b] == 0xFFFFFFFFu ? elems : c->bu_count[b];
    }
    return c->total ? (double)keys / (double)c->total : 1.0;
}

// Called at the end of an upload.  TCLUST_REORDER=0 never builds the copy,
// =1 always uses it; default: build it when the input order leaves the union lists above 60 % of the elements (the
// point below which the block join pays, kj_unit_pays) and use it when its lists are at least a fifth shorter.
int tc_shadow_build(tc_ctx *ctx)
{
    ctx->use_shadow = false;
    const char *env = getenv("TCLUST_REORDER");
    const bool never = env && !strcmp(env, "0"), always = env && !strcmp(env, "1");
    if (never || ctx->total == 0 || ctx->n < 64 || !ctx->d_multi) return 0;
    const double own = list_ratio(ctx);
    if (!always && (ctx->n < 2048 || own <= 0.6)) return 0;
    const uint32_t n = ctx->n;
    cudaStream_t st = ctx->stream;
    if (!ctx->shadow) {
        ctx->shadow = new tc_ctx();
        ctx->shadow->device = ctx->device; ctx->shadow->num_sms = ctx->num_sms; ctx->shadow->stream = ctx->stream;
        ctx->shadow->ev_pass0 = ctx->ev_pass0; ctx->shadow->ev_pass1 = ctx->ev_pass1;
        ctx->shadow->d_edge_count = ctx->d_edge_count;
    }
    tc_ctx *s = ctx->shadow;
    s->kh_edge_units.clear(); s->kh_edge_units_n = 0; s->kh_units_on_device = false; s->kj_split_valid = false;
    s->n = n; s->total = ctx->total;
    s->d_multi = ctx->d_multi;                                  // same ranks, same bitmap (owned by ctx)
    // ---- permutation: stable sort of the sketch ids by locality key
    uint32_t *d_key = nullptr;                                  // key in, key out, id in | id out = d_perm
    size_t tmp_bytes = 0;
    TC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                            (uint32_t *)nullptr, (int)n, 0, 32, st));
    if (grow(&s->d_perm, &s->cap_perm, n)) return 1;
    {   // the shadow's own scratch (the owner's still holds d_scan)
        void *ws = nullptr;
        if (tc_scratch(s, (size_t)n * 12 + tmp_bytes + 256, &ws)) return 1;
        d_key = static_cast<uint32_t *>(ws);
    }
    uint32_t *d_key_out = d_key + n, *d_idx = d_key + 2 * (size_t)n;
    void *d_tmp = reinterpret_cast<uint8_t *>(d_key) + (((size_t)n * 12 + 255) & ~(size_t)255);
    k3r_key_kernel<<<(n + 127) / 128, 128, 0, st>>>(ctx->d_ranks, reinterpret_cast<unsigned long long *>(ctx->d_off), ctx->d_multi,
                                                   n, d_key, d_idx);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key_out, d_idx, s->d_perm, (int)n, 0, 32, st);
    std::vector<uint32_t> perm(n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(perm.data(), s->d_perm, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    ctx->launches += 4;
    if (e != cudaSuccess) { tc_set_error("tc_shadow_build: %s", cudaGetErrorString(e)); return 1; }
    // ---- the copy: offsets on the host, ranks gathered on the device
    s->h_size.resize(n); s->h_off.resize((size_t)n + 1);
    s->h_off[0] = 0;
    for (uint32_t i = 0; i < n; i++) {
        s->h_size[i] = ctx->h_size[perm[i]];
        s->h_off[i + 1] = s->h_off[i] + s->h_size[i];
    }
    if (grow(&s->d_ranks, &s->cap_ranks, ctx->total)) return 1;
    if (n + 1 > s->cap_n) {
        cudaFree(s->d_off); cudaFree(s->d_size);
        s->d_off = nullptr; s->d_size = nullptr; s->cap_n = 0;
        TC_CUDA(cudaMalloc(&s->d_off, (size_t)(n + 1) * 8));
        TC_CUDA(cudaMalloc(&s->d_size, (size_t)(n + 1) * 4));
        s->cap_n = n + 1;
    }
    TC_CUDA(cudaMemcpyAsync(s->d_off, s->h_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    k3r_gather_kernel<<<n, 128, 0, st>>>(ctx->d_ranks, reinterpret_cast<unsigned long long *>(ctx->d_off), s->d_perm,
                                         reinterpret_cast<unsigned long long *>(s->d_off), n, s->d_ranks, s->d_size);
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(st));
    if (tc_kh_prepare(s)) return 1;
    ctx->launches += s->launches; s->launches = 0;
    const double theirs = list_ratio(s);
    ctx->use_shadow = always || theirs <= 0.8 * own;
    if (getenv("TCLUST_TIMING"))
        fprintf(stderr, "[tclust timing] union lists keep %.3f of the elements in input order, %.3f reordered: %s\n", own, theirs,
                ctx->use_shadow ? "dense pass on the reordered copy" : "dense pass in input order");
    return 0;
}

// the dense pass (all pairs, semantic A) on the reordered copy; edges land in the owner's buffer with the owner's ids
int tc_shadow_pass(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, bool force_block)
{
    tc_ctx *s = ctx->shadow;
    s->d_edges = ctx->d_edges; s->edge_cap = ctx->edge_cap;
    s->d_kh_state = ctx->d_kh_state;
    const int rc = tc_launch_blockjoin_edges(s, radio, part, nparts, 0, force_block);
    ctx->d_kh_state = s->d_kh_state;
    ctx->launches += s->launches; s->launches = 0;
    if (rc) return rc;
    return tc_shadow_remap_from(ctx, 0);
}

int tc_shadow_remap_from(tc_ctx *ctx, uint64_t first_edge)
{
    tc_ctx *s = ctx->shadow;
    k3r_remap_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->d_edges, ctx->d_edge_count, ctx->edge_cap, s->d_perm, first_edge);
    if (ctx->ev_pass1) cudaEventRecord(ctx->ev_pass1, ctx->stream);      // tc_last_pass_kernel_ms covers the remap too
    ctx->launches++;
    TC_CUDA(cudaGetLastError());
    return 0;
}
