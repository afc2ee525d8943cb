// tc_common.cuh -- context object and helpers shared by the kernels of libtclust_cuda.so
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/tclust.h"

void tc_set_error(const char *fmt, ...);

#define TC_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            tc_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

#define TC_CHECK(cond, ...)                                                            \
    do {                                                                               \
        if (!(cond)) { tc_set_error(__VA_ARGS__); return 1; }                          \
    } while (0)

constexpr uint32_t TC_GROUP = 32;            // sketches per interleaved group (= warp width)
constexpr int TC_UP_CHUNKS = 8;             // pieces of a chunked host upload (tclust_api.cu, upload_sorted_runs)
constexpr uint32_t TC_SENT_A = 0xFFFFFFFFu;  // end-of-window sentinel, row streams
constexpr uint32_t TC_SENT_B = 0xFFFFFFFEu;  // end-of-window sentinel, column streams
constexpr uint32_t TC_RANK_MAX = 0xFFFFFFFDu;

struct tc_kh_unit { uint32_t rb, c0, c1; };   // K3H work unit: row block x column segment [c0, c1)

// K4 (k4_boruvka.cu): device-resident Boruvka state; xbuf is the buffer peers map in a multi-GPU group
struct tc_k4_state {
    unsigned long long *d_xbuf = nullptr; uint64_t cap_xbuf = 0; uint32_t xbuf_n = 0;
    unsigned long long *d_red = nullptr;  uint64_t cap_red = 0;
    uint32_t *d_parent = nullptr;         uint64_t cap_parent = 0;
    unsigned long long *d_out = nullptr;  uint64_t cap_out = 0;      // forest (key, id) + sort alternates: 4 x n
    unsigned long long *d_fin = nullptr;  uint64_t cap_fin = 0;      // forest of the edge batches seen so far (bounded-memory MST): 2 x n
    uint32_t *d_ctl = nullptr;
    unsigned char *d_live = nullptr;      uint64_t cap_live = 0;     // liveness flag per chunk of candidate edges
    void *d_sort_tmp = nullptr;           uint64_t cap_sort_tmp = 0;
    void *h_out = nullptr;                uint64_t h_out_bytes = 0;  // page-locked read-back buffer
    int grid = 0;
    int rank = 0, nranks = 0;             // multi-GPU group (0/1: single GPU)
    unsigned long long *peer_xbuf[8] = {};
    bool peer_opened[8] = {};             // mapped with cudaIpcOpenMemHandle (another process), to be closed
    unsigned long long sys_base = 0;
    bool export_fresh = false;            // tc_mg_export done and not yet consumed by a tc_mg_join
    uint32_t last_rounds = 0;             // Boruvka rounds of the last tc_mst
    cudaEvent_t ev[3] = {};               // before / after the Boruvka kernel, after the forest sort
};

struct tc_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    uint64_t last_h2d_bytes = 0;          // sequence bytes the last tc_sketch_minhash moved over PCIe (packed batches count packed)
    cudaEvent_t ev_pass0 = nullptr, ev_pass1 = nullptr;   // around the last dense-pass kernel
    uint32_t attr_set = 0;                // dynamic shared-memory opt-ins done on this context's device (bit per file)
    bool boruvka_async = false;           // tc_boruvka_set_async: the Boruvka phases do not synchronise the stream

    // resident sketches (HP-2)
    uint32_t n = 0;               // number of sketches
    uint64_t total = 0;           // total hashes
    uint64_t *d_hashes = nullptr; // original u64 hashes, ragged row-major (independent validation kernel only)
    bool hashes_resident = false; // false after an upload from sorted runs (tc_sketches_upload_runs_dev)
    uint64_t *d_off = nullptr;    // n+1 offsets
    uint32_t *d_size = nullptr;   // n sizes
    uint32_t *d_ranks = nullptr;  // order-preserving dense u32 ranks of the hashes, same layout as d_hashes
    uint32_t ngroups = 0;
    bool gt_valid = false;        // d_gt / d_goff hold the current set (built on first use by the merge kernel's launchers)
    uint32_t *d_gt = nullptr;     // group-interleaved ranks: [group][row][lane]
    uint64_t *d_goff = nullptr;   // ngroups+1 element offsets into d_gt
    uint64_t cap_hashes = 0, cap_ranks = 0, cap_gt = 0, cap_n = 0, cap_groups = 0, cap_comp = 0;   // element capacities
    std::vector<uint64_t> h_off;
    std::vector<uint32_t> h_size;

    // candidate edges (semantic-A graph) produced by the dense pass
    tc_pair *d_edges = nullptr;
    uint64_t edge_cap = 0;
    uint32_t edge_cap_for_n = 0;  // n the edge buffer was last sized for (0: re-evaluate)
    unsigned long long *d_edge_count = nullptr;  // [0] = count, [1] = overflow flag, [2] = tile counter
    uint64_t n_edges = 0;
    uint64_t bounded_batches = 0, bounded_folds = 0;   // last bounded-memory MST (0: the candidate graph fitted)
    bool edges_partial = false;       // d_edges holds only part of the candidate graph (bounded-memory MST, one rank of a group)
    uint32_t edges_semb_size = 0;     // > 0: the edges carry semantic-B counts taken with this sketchSize
    int32_t *d_comp = nullptr;

    // representative table of the greedy scan (K5): same interleaved layout, members in creation order
    uint32_t nrep = 0;
    uint32_t *d_rep_gt = nullptr;   uint64_t rep_gt_cap = 0;
    uint64_t *d_rep_goff = nullptr; uint64_t rep_goff_cap = 0;
    uint32_t *d_rep_len = nullptr;  uint64_t rep_len_cap = 0;
    uint32_t *d_rep_ids = nullptr;  uint64_t rep_ids_cap = 0;
    std::vector<uint32_t> h_rep_ids;
    std::vector<uint64_t> h_rep_goff = std::vector<uint64_t>(1, 0);
    tc_cand *d_cands = nullptr;
    uint64_t cand_cap = 0;

    // hash-join pass (K3H): partition pivots per row block, unit list, per-CTA column state
    uint32_t *d_kh_pivots = nullptr, *d_kh_pivoff = nullptr, *d_kh_state = nullptr;
    tc_kh_unit *d_kh_units = nullptr;
    size_t kh_units_cap = 0;
    std::vector<tc_kh_unit> kh_edge_units;
    uint32_t kh_edge_units_n = 0, kh_edge_units_start = 0, kh_edge_units_cseg = 0;
    bool kh_units_on_device = false;     // d_kh_units currently holds kh_edge_units
    // K3J block join: per 32-sketch block the sorted distinct ranks + 32-bit membership masks, stored at the element
    // offset of the block's first sketch; bu_count[b] = number of distinct ranks (0xFFFFFFFF: block too large, no list)
    uint32_t *d_bu_keys = nullptr, *d_bu_mask = nullptr, *d_bu_count = nullptr;
    uint32_t *d_multi = nullptr;         // bitmap over the ranks: bit r set = rank r occurs in at least two sketches
    uint64_t cap_multi = 0;              // words
    uint64_t cap_bu = 0, cap_bu_count = 0, cap_pivoff = 0, cap_pivots = 0, kj_units_cap = 0;   // grow-only: no cudaMalloc/cudaFree per upload
    std::vector<uint32_t> bu_count;
    std::vector<uint32_t> kh_parts;      // table partitions per row block
    std::vector<tc_kh_unit> kj_units_blk, kj_units_old;   // this rank's units routed to K3J / K3H, resident in d_kj_units
    tc_kh_unit *d_kj_units = nullptr;
    // K3S (sparse rows): their units, the flagged tiles as a device list, the list length (device / host copy)
    std::vector<tc_kh_unit> ks_units;
    void *d_ks_tiles = nullptr; uint64_t cap_ks_tiles = 0;
    uint32_t *d_ks_count = nullptr; uint32_t ks_ntiles_host = 0;
    uint64_t kj_split_key = 0;
    bool kj_split_valid = false;
    // tile candidacy (k3t_tileflags.cu): bit (rb, cb) of row rb = some rank occurs both in block rb and in block cb
    uint32_t *d_tile_flags = nullptr, *d_tile_rowcount = nullptr;
    uint64_t cap_tile_flags = 0, cap_tile_rowcount = 0;
    uint32_t tile_words = 0;             // words per row of d_tile_flags (0: no flags, every tile is visited)
    std::vector<uint32_t> tile_rowcount; // flagged tiles per row block
    uint64_t tiles_flagged = 0;
    uint64_t pairs_cache_key = 0, pairs_cache_val = 0;     // n_pairs of the last sharded pass (tc_allpairs_pass, nparts > 1)
    bool pairs_cache_valid = false;

    // reordered shadow copy of the sketch set for the dense pass (k3r_reorder.cu); d_perm (held by the shadow): new id -> id
    tc_ctx *shadow = nullptr;
    bool use_shadow = false;
    uint32_t *d_perm = nullptr;
    uint64_t cap_perm = 0;

    // KSSD shuffled_dim table (K2), resident once uploaded
    int32_t *d_kssd_dim = nullptr;
    uint32_t *d_kssd_filter = nullptr;   // presence bitmap of the passing slots (k2_kssd.cu)
    int kssd_dim_end = -1;
    const int32_t *kssd_dim_host = nullptr;
    size_t kssd_dim_n = 0;

    // K1 pipeline workspace (k1_sketch.cu): staging slots, copy stream
    void *k1_ws = nullptr;
    void *k2_ws = nullptr;       // K2 (KSSD) pipeline workspace (k2_kssd.cu)

    // page-locked staging of tc_sketches_upload_rows (two buffers, allocated on first use)
    uint64_t *h_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    // tc_sketches_upload of a large host array: copy stream + one event per piece (+1: "work stream idle")
    cudaStream_t up_copy_stream = nullptr;
    cudaEvent_t ev_chunk[TC_UP_CHUNKS + 1] = {};

    tc_k4_state k4;
    // threshold-only clustering: while set, the edge passes union (i, j) on the device instead of appending edges
    uint32_t *uf_parent = nullptr; const uint32_t *uf_perm = nullptr; double uf_jmin = 2.0; int uf_containment = 0;
    uint32_t *d_uf = nullptr; uint64_t cap_uf = 0;
    bool pass_nosync = false;            // dense-pass launchers do not wait for the device (tc_mst queues pass + Boruvka back to back)

    // scratch
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
};

int tc_scratch(tc_ctx *ctx, size_t bytes, void **out);
int tc_ensure_interleaved(tc_ctx *ctx);    // tclust_api.cu
void tc_k1_free(tc_ctx *ctx);
// device Boruvka (k4_boruvka.cu)
size_t tc_k4_xbuf_words(uint32_t n);
void tc_k4_free(tc_ctx *ctx);
int tc_k4_reserve(tc_ctx *ctx, uint32_t n);
int tc_k4_enqueue(tc_ctx *ctx, int mode, const unsigned long long *d_fin_key, const unsigned long long *d_fin_id, uint32_t nfin);
int tc_k4_collect(tc_ctx *ctx, int mode, int kmer_size, std::vector<tc_edge> &mst, uint64_t *n_edges_dev, bool *overflow,
                  std::vector<int32_t> *comp_out, tc_edge *direct = nullptr, uint64_t *n_direct = nullptr);
int tc_k4_fold(tc_ctx *ctx, uint32_t *nfin);     // wait for the queued Boruvka and keep its forest on the device as the next launch's extra edges
void tc_k2_free(tc_ctx *ctx);

// tile-kernel entry points (k3_tile.cu)
// semb_size > 0 selects semantic B (Mash union-limited merge, MinHash.cpp:982-1101) with that sketchSize
int tc_launch_tiles_matrix(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *d_out,
                           uint32_t semb_size = 0);
int tc_launch_tiles_edges(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, uint32_t semb_size = 0);
int tc_launch_tiles_rep_matrix(tc_ctx *ctx, uint32_t q0, uint32_t q1, uint32_t *d_common, uint32_t *d_first_hit);
int tc_launch_tiles_cands(tc_ctx *ctx, uint32_t q0, uint32_t q1, int vs_reps, int cmin_fixed, int containment,
                          double jmin, uint32_t part, uint32_t nparts, uint32_t semb_size = 0);
// hash-join pass (k3h_join.cu)
int tc_kh_prepare(tc_ctx *ctx);
int tc_tile_flags_build(tc_ctx *ctx, const std::vector<uint32_t> &dist);   // k3t_tileflags.cu
int tc_launch_join_matrix(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *d_out);
int tc_launch_blockjoin_matrix(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *d_out);
int tc_launch_blockjoin_edges(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, uint32_t start_index, bool force_block,
                              bool plan_only = false);
size_t tc_blockjoin_list_size(const tc_ctx *ctx, int list);            // list 0 = sparse tiles (K3S), 1 = K3J units, 2 = K3H units
int tc_blockjoin_run_range(tc_ctx *ctx, int radio, uint32_t start_index, int list, size_t u0, size_t u1);
uint64_t tc_blockjoin_unit_bound(const tc_ctx *ctx, int list, size_t u);
int tc_launch_join_cands(tc_ctx *ctx, uint32_t q0, uint32_t q1, int vs_reps, int cmin_fixed, int containment, double jmin);
int tc_launch_join_edges(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, uint32_t start_index = 0);
// reordered shadow copy (k3r_reorder.cu)
int tc_shadow_build(tc_ctx *ctx);
int tc_shadow_pass(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts, bool force_block);
int tc_shadow_remap_from(tc_ctx *ctx, uint64_t first_edge);      // candidate edges [first_edge, count): ids of the copy -> input ids
void tc_shadow_free(tc_ctx *ctx);
// representative table (k5_greedy.cu)
void tc_rep_table_free(tc_ctx *ctx);
void tc_rep_table_reset(tc_ctx *ctx);
int tc_rep_table_append(tc_ctx *ctx, const uint32_t *ids, uint32_t count);
int tc_launch_pair_common_ref(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1, uint32_t *d_out);
