/*
 * tclust.h -- C ABI of libtclust_cuda.so: the B200 (sm_100a) implementation of the two
 * data-parallel hot paths of RabbitTClust.
 *
 * The reference has no FFI layer; the hot paths sit behind the library API its callers
 * use.  Each entry point below names the reference interface it stands
 * in for (paths relative to the reference root).  INTEGRATION.md shows the call sites a
 * maintainer would change.
 *
 * Conventions: every function returns 0 on success, non-zero on error (message via
 * tc_last_error(), thread-local).  The caller owns all host buffers.  Device memory is
 * owned by the opaque tc_ctx (one per process per GPU, single caller thread).  Pointers
 * named *_dev are CUDA device pointers on the context's device; all others are host
 * pointers.  There is NO CPU fallback: without a CUDA device tc_ctx_create fails.
 *
 * Results are integers (hashes, intersection sizes, ids).  Distances are produced on the
 * host by tc_*_dist below with the reference's exact expression order and glibc log(), so
 * they are bit-identical to the reference's doubles.
 */
#ifndef TCLUST_H
#define TCLUST_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TC_VERSION 100
typedef struct tc_ctx tc_ctx;

/* src/MST.h:17-21 EdgeInfo {int preNode; int sufNode; double dist;} (edge.mst record,
 * src/MST_IO.cpp:200-217). */
typedef struct { int32_t pre, suf; double dist; } tc_edge;

/* one candidate edge of the semantic-A graph: pair (i, j<i) with common >= 1 */
typedef struct { uint32_t i, j, common; } tc_pair;

/* one candidate of the greedy representative scan: query genome, representative genome, shared
 * hashes, and the index in the query sketch of the first shared hash (what greedy.cpp:1148-1164
 * accumulates in cnt[] / the position of the representative in touched[]) */
typedef struct { uint32_t query, rep, common, first_hit; } tc_cand;

int         tc_version(void);
const char *tc_last_error(void);

/* ---- context ------------------------------------------------------------------------ */
/* The reference keeps all state in host vectors owned by its callers (vector<SketchInfo>, SketchInfo.h:23-36) and
 * has nothing to create; the context is what SURVEY.md 8b-3 introduces for the device-resident copies: one per
 * device, one caller thread at a time.  Several devices: one context each, joined by tc_mg_join (below). */
int  tc_ctx_create(int device, tc_ctx **out);
void tc_ctx_destroy(tc_ctx *ctx);
/* measurement aids without a counterpart in the reference (its Timer prints, src/sub_command.cpp `#ifdef Timer`, are
 * what the CLI reproduces): number of kernels of this library launched on the context so far (bench accounting) */
uint64_t tc_kernel_launches(const tc_ctx *ctx);
/* duration of the last dense-pass kernel (tc_allpairs_pass*, tc_mst*), CUDA events recorded on the context's stream
 * immediately around the launch; -1 if no pass has completed */
double tc_last_pass_kernel_ms(const tc_ctx *ctx);
/* CUDA stream (cudaStream_t) the context launches on, for event timing by the caller */
void *tc_stream(const tc_ctx *ctx);

/* page-locked host memory for sequence buffers handed to tc_sketch_* (the FASTA feeder parses straight into
 * it): copies from it overlap the kernels; pageable buffers work too, only slower. */
void *tc_host_alloc(size_t bytes);
void  tc_host_free(void *p);

/* ---- HP-1: sketch construction ------------------------------------------------------- */
/*
 * MinHash bottom-s sketches of `ngenomes` genomes.  Replaces, per genome,
 *   Sketch::MinHash mh(k, size, seed=42); for each record: mh.update(seq); mh.storeMinHashes()
 * (RabbitSketch/src/Sketch.h:60-101, MinHash.cpp:465-747, 287-348, 756-843; driver
 * src/SketchInfo.cpp:918-948).
 * seqs: concatenated record bytes (any case, any byte; upper-cased on device as
 * MinHash.cpp:472-479); record r = seqs[rec_off[r] .. rec_off[r+1]); rec2genome[r] = genome
 * of record r (non-decreasing).  k <= 32.  sketch_size[g] = s of genome g.
 * out_hashes: ngenomes rows of out_stride u64 (out_stride >= max sketch_size); row g holds
 * out_n[g] <= sketch_size[g] hashes ascending (the s smallest distinct).
 * The call pipelines batches of whole genomes (about TCLUST_SKETCH_CHUNK_MB = 128 MB of sequence each) over
 * two device staging buffers: the H2D copy of batch c+1 overlaps the kernels of batch c.
 */
int tc_sketch_minhash(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off,
                      const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes, int k,
                      const uint32_t *sketch_size, uint32_t seed,
                      uint64_t *out_hashes, uint64_t out_stride, uint32_t *out_n);
/* Transport: when the host has cores to spare (>= 12 threads per visible GPU; TCLUST_SKETCH_PACK=0|1 forces it off/on,
 * TCLUST_PACK_THREADS sets the thread count) every batch is packed on the host to 2 bits per base plus an exception list
 * of the bytes that are not ACGTacgt, a quarter of the bytes crosses PCIe, and the device rebuilds the exact byte stream
 * before sketching -- same hashes, bit for bit.  tc_last_h2d_bytes: sequence bytes the last call moved over PCIe. */
uint64_t tc_last_h2d_bytes(const tc_ctx *ctx);
/* same with all buffers already resident on the device (timed "value" leg of bench.py);
 * seqs_dev must be readable 16 bytes past its end. */
int tc_sketch_minhash_dev(tc_ctx *ctx, const uint8_t *seqs_dev, const uint64_t *rec_off,
                          const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes, int k,
                          const uint32_t *sketch_size, uint32_t seed,
                          uint64_t *out_hashes_dev, uint64_t out_stride, uint32_t *out_n_dev);

/*
 * KSSD (--fast) sketches, list mode.  Replaces the hot loop of sketchFileWithKssd
 * (src/SketchInfo.cpp:1019-1048 constants, 1126-1165 loop, 1180-1193 sort).
 * shuffled_dim: host table of 1<<(4*half_subk) ints (generate_shuffle_dim,
 * SketchInfo.cpp:60-102; glibc rand() -- host-only by construction; tc_kssd_shuffled_dim
 * builds it, half_subk = 6 unless drlevel > 4, SketchInfo.cpp:1022).
 * out: ngenomes rows of out_stride u64, ascending, distinct; out_n[g] = count.  If a genome yields more than out_stride
 * hashes the call returns TC_ERR_KSSD_STRIDE (nothing is truncated silently).
 */
#define TC_ERR_KSSD_STRIDE 2   /* tc_sketch_kssd*: some genome needs a longer output row; tc_kssd_needed_stride() says how long */
int tc_kssd_shuffled_dim(int half_subk, int32_t *out /* 1 << (4*half_subk) ints */);
/* after tc_sketch_kssd* returned TC_ERR_KSSD_STRIDE on this thread: the largest number of hashes (before deduplication)
 * any genome of that call produced -- retry with out_stride >= it */
uint32_t tc_kssd_needed_stride(void);
int tc_sketch_kssd(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off,
                   const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes,
                   int kmer_size, int drlevel, const int32_t *shuffled_dim,
                   uint64_t *out_hashes, uint64_t out_stride, uint32_t *out_n);
/* One KSSD sketch per RECORD of a multi-FASTA: the single-file consumer consumer_fasta_task_with_kssd
 * (src/SketchInfo.cpp:274-433; caller sketchSequencesWithKssd, :554-642).  nrec == ngenomes, rec2genome[r] == r.  For an even
 * kmer_size the hash sets are those of tc_sketch_kssd.  For an ODD one the reference does not round kmerSize up to
 * 2*half_k here (:296 vs :1019-1020): a window is used after kmer_size valid bases while the rolling tuple holds
 * kmer_size+1, a non-ACGT byte restarts the count without clearing the tuple (:368-371), so the first window after every
 * gap keeps the last valid base from before it (nothing at the start of a record), and the outer-context bits are packed
 * with the un-rounded size (:386).  All of that is reproduced on the device.  The reference leaves each set in hash-set
 * iteration order; rows come back ascending. */
int tc_sketch_kssd_seq(tc_ctx *ctx, const uint8_t *seqs, const uint64_t *rec_off,
                       const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes,
                       int kmer_size, int drlevel, const int32_t *shuffled_dim,
                       uint64_t *out_hashes, uint64_t out_stride, uint32_t *out_n);
/* same with sequence and outputs resident on the device (shuffled_dim stays a host pointer; the
 * table is uploaded once per context and pointer) */
int tc_sketch_kssd_dev(tc_ctx *ctx, const uint8_t *seqs_dev, const uint64_t *rec_off,
                       const uint32_t *rec2genome, uint64_t nrec, uint32_t ngenomes,
                       int kmer_size, int drlevel, const int32_t *shuffled_dim,
                       uint64_t *out_hashes_dev, uint64_t out_stride, uint32_t *out_n_dev);

/* ---- HP-2: all-pairs sorted-sketch intersection -------------------------------------- */
/*
 * Make N sorted, duplicate-free sketches resident on the device (what compute_minhash_mst
 * caches at src/MST.cpp:1311-1322 and greedy.cpp:1138).  sketch g = hashes[off[g]..off[g+1]).
 * Replaces any previously uploaded set.  At most 2^31 - 1 hashes in total per context (an error otherwise, never a
 * truncation); a sketch that is not strictly ascending is an error naming the sketch.
 */
int tc_sketches_upload(tc_ctx *ctx, const uint64_t *hashes, const uint64_t *off, uint32_t n);
int tc_sketches_upload_dev(tc_ctx *ctx, const uint64_t *hashes_dev, const uint64_t *off, uint32_t n);
/* the same set given the way the reference holds it -- one array per sketch (vector<SketchInfo>, each with its own
 * storeMinHashes() / hash64_arr vector): rows[g] points at sizes[g] hashes.  The rows are gathered through page-locked
 * staging buffers while earlier chunks are already on their way to the device, so the caller needs no flat copy. */
int tc_sketches_upload_rows(tc_ctx *ctx, const uint64_t *const *rows, const uint64_t *sizes, uint32_t n);

/*
 * Multi-GPU upload ("every GPU holds all N sketches for HP-2"; the reference is a single-node OpenMP
 * program and has no counterpart -- its threads share the one vector<SketchInfo>).  Every GPU of a group needs all sketches
 * resident, but each should move and sort only its share:
 * rank r copies its contiguous share [pos_base, pos_base + len) of the flat hash array to its device and calls
 * tc_upload_sort_share_dev, which sorts the share by hash and pairs every hash with its position in the flat array; the
 * ranks all-gather the sorted runs (keys + positions; NCCL, plumbing) and every rank calls tc_sketches_upload_runs_dev,
 * which merges the runs instead of sorting everything again and then proceeds exactly like tc_sketches_upload.  The
 * original u64 hashes are not kept (tc_pair_common impl 1 is unavailable after it).  Used by bench.py (e2e, N > 1).
 */
int tc_upload_sort_share_dev(tc_ctx *ctx, const uint64_t *hashes_dev, uint64_t len, uint64_t pos_base,
                             uint64_t *keys_out_dev, uint32_t *pos_out_dev);
int tc_sketches_upload_runs_dev(tc_ctx *ctx, const uint64_t *keys_dev, const uint32_t *pos_dev, const uint64_t *run_off,
                                uint32_t nruns, const uint64_t *off, uint32_t n);

/*
 * Multi-GPU upload, second form: the rank transform itself is sharded by hash VALUE RANGE, so that no rank sorts more than
 * its 1/G of the hashes and only 4-byte ranks (not 12-byte sorted runs) are gathered.  Per rank (exchanges = the caller's
 * all-to-all / all-gather; rabbittclust_b200/mst_dist.py:upload_sharded_vr is the orchestration bench.py uses):
 *   1. copy the own contiguous share of the flat hash array to the device; agree on G-1 splitters (a sorted sample);
 *   2. tc_vr_partition_dev(mode 0): (hash, position) pairs grouped by value range -> all-to-all;
 *   3. tc_vr_local_ranks_dev: sort the received pairs, dense local ranks + "occurs more than once" flag (bit 31);
 *      all-gather the distinct counts -> base of every range;
 *   4. tc_vr_partition32_dev: (rank, position) grouped by the rank that owns the position -> all-to-all back;
 *   5. tc_vr_scatter_ranks_dev: own slice of the position-ordered rank array (base of the source range added);
 *      all-gather the slices;
 *   6. tc_sketches_upload_ranks_dev: resident set from the finished ranks (as tc_sketches_upload from there on).
 * Same resident state as tc_sketches_upload (identical ranks, since the ranges are ordered and ranks are dense).
 */
int tc_vr_partition_dev(tc_ctx *ctx, const uint64_t *a_dev, const uint32_t *b_dev /* NULL: pos_base + index */, uint64_t m, int mode,
                        const uint64_t *splitters /* host, nb - 1, ascending (mode 0) */, uint64_t per /* mode 1 */, uint32_t nb,
                        uint32_t pos_base, uint64_t *a_out_dev, uint32_t *b_out_dev, uint64_t *counts /* host, nb */);
int tc_vr_partition32_dev(tc_ctx *ctx, const uint32_t *a_dev, const uint32_t *b_dev, uint64_t m, uint64_t per, uint32_t nb,
                          uint32_t *a_out_dev, uint32_t *b_out_dev, uint64_t *counts);
int tc_vr_local_ranks_dev(tc_ctx *ctx, const uint64_t *keys_dev, const uint32_t *pos_dev, uint64_t m,
                          uint32_t *pos_out_dev, uint32_t *rank_out_dev, uint64_t *n_distinct);
int tc_vr_scatter_ranks_dev(tc_ctx *ctx, const uint32_t *pos_dev, const uint32_t *rank_dev, const uint64_t *seg_off /* host, nseg + 1 */,
                            const uint64_t *seg_base /* host, nseg */, uint32_t nseg, uint32_t pos_lo, uint64_t share_len,
                            uint32_t *out_dev);
int tc_sketches_upload_ranks_dev(tc_ctx *ctx, const uint32_t *ranks_flag_dev, const uint64_t *off, uint32_t n);

/* validation: common[i-i0][j-j0] = |sketch_i n sketch_j| for i in [i0,i1), j in [j0,j1)
 * (semantic A: the counts of src/MST.cpp:1409-1435).  impl 0 = tiled merge kernel (k3_tile.cu),
 * 1 = independent u64 global-memory merge kernel, 2 = hash-join kernel (k3h_join.cu). */
int tc_pair_common(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1,
                   int impl, uint32_t *out);

/* semantic B validation: out[i-i0][j-j0] = number of shared hashes among the first `sketch_size` elements of the
 * sorted union of sketch_i and sketch_j -- the `common` of Sketch::MinHash::jaccard
 * (RabbitSketch/src/MinHash.cpp:982-1101; stop rule u64_intersect_scalar_stop :1199-1229, to which the AVX2
 * kernel :1579-1643 is equivalent).  Its denominator is min(sketch_size, |A|+|B|-common) (tc_jaccard_b). */
int tc_pair_common_b(tc_ctx *ctx, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1,
                     uint32_t sketch_size, uint32_t *out);

/*
 * One dense pass over all pairs j < i of the uploaded sketches, restricted to the tiles
 * owned by `part` of `nparts` (multi-GPU sharding; 0/1 = everything).  Emits the candidate
 * edges of compute_minhash_mst: common >= 1 and max(size) <= radio*min(size)
 * (src/MST.cpp:1470,1481-1487).  Edges stay on the device for tc_mst_*; n_pairs / n_edges
 * (may be NULL) receive the pairs evaluated and edges kept by this call.  The edge list has no defined order; (i, j)
 * are always ids of the uploaded set with j < i, also when the pass ran on the library's internally reordered copy.
 */
int tc_allpairs_pass(tc_ctx *ctx, int radio, uint32_t part, uint32_t nparts,
                     uint64_t *n_pairs, uint64_t *n_edges);
/* the append form (start_index of compute_minhash_mst, src/MST.cpp:1375-1383): only rows i >= start_index, i.e. the
 * pairs (new genome, any earlier genome) */
int tc_allpairs_pass_from(tc_ctx *ctx, int radio, uint32_t start_index, uint32_t part, uint32_t nparts,
                          uint64_t *n_pairs, uint64_t *n_edges);
/* the same pass with semantic-B counts (the all-pairs loop of modifyMST, src/MST.cpp:851-868): edges are the
 * pairs with common_B >= 1, no size-ratio filter; the remaining pairs have distance exactly 1. */
int tc_allpairs_pass_b(tc_ctx *ctx, uint32_t sketch_size, uint32_t part, uint32_t nparts,
                       uint64_t *n_pairs, uint64_t *n_edges);
/* copy the candidate edges kept so far to the host (cap entries at most; *n = total) */
int tc_edges_download(tc_ctx *ctx, tc_pair *out, uint64_t cap, uint64_t *n);

/*
 * Boruvka over the device-resident candidate edges.  One round: per current component
 * (comp[v], host array of N labels) the minimum-weight outgoing edge under the total order
 * (distance asc == similarity key desc, then (i,j) asc).  Two reductions per round so that
 * ranks can all-reduce(min) between them:
 *   phase 1: best_key[c]  = min over outgoing edges of key      (N u64, device buffer)
 *   phase 2: best_edge[c] = min (i<<32|j) among edges whose key == best_key[c]
 * key = ~bits(J) with J = (double)common/denom exactly as src/MST.cpp:1489-1492 (Mash) or
 * :1505-1507 (containment) computes it, so the host can turn it into the reference's
 * distance.  The *_dev buffers are caller-provided (e.g. torch tensors that NCCL reduces).
 * `containment` is the distance mode TC_DIST_*: 0/1 as above; TC_DIST_MASH_B after tc_allpairs_pass_b
 * (J = common_B / min(sketchSize, union)); TC_DIST_CONTAIN_B = containDistance (J = common / min size).
 */
#define TC_DIST_MASH_A    0   /* src/MST.cpp:1489-1503 */
#define TC_DIST_AAF       1   /* src/MST.cpp:1504-1515 */
#define TC_DIST_MASH_B    2   /* RabbitSketch/src/MinHash.cpp:1125-1153 MinHash::distance */
#define TC_DIST_CONTAIN_B 3   /* RabbitSketch/src/MinHash.cpp:1102-1123 MinHash::containDistance */
int tc_boruvka_upload_comp(tc_ctx *ctx, const int32_t *comp, uint32_t n);
int tc_boruvka_phase1(tc_ctx *ctx, int containment, uint64_t *best_key_dev);
int tc_boruvka_phase2(tc_ctx *ctx, int containment, const uint64_t *best_key_dev,
                      uint64_t *best_edge_dev);
/* on = 1: tc_boruvka_upload_comp / _phase1 / _phase2 only enqueue their work on tc_stream(ctx) and return; the
 * caller orders what follows (the NCCL all-reduce of the *_dev buffers, the next phase) on that stream and
 * synchronises once per round.  on = 0 (default): every phase returns with its result complete.  `comp` must stay
 * valid until the stream has consumed it. */
int tc_boruvka_set_async(tc_ctx *ctx, int on);

/*
 * Whole single-GPU MST build = compute_minhash_mst (src/MST.cpp:1290-1737): dense pass +
 * Boruvka rounds + host distances.  out: up to n-1 forest edges {i, j<i, dist} sorted by
 * (dist, i, j); *n_out = count.  The sorted multiset of weights and the partition at any
 * threshold equal the reference's.
 */
int tc_mst(tc_ctx *ctx, int kmer_size, int containment, double threshold,
           tc_edge *out, uint64_t *n_out);

/*
 * Threshold-only clustering: the clusters of clust-mst at `threshold` without the MST -- connected
 * components of the candidate pairs whose distance is <= threshold, i.e. the partition generateForest +
 * generateClusterWithBfs (src/MST.cpp:77-142) cut out of the MST.  The dense pass unions the pairs in a device-resident
 * union-find instead of writing edges: O(n) memory whatever the candidate graph, no Boruvka.  labels[v] = smallest id
 * of v's cluster, over the pairs of `part` of `nparts` (multi-GPU: one part per GPU, then tc_labels_merge on the host).
 * The within-threshold test is exact: J = common/denom against the smallest J whose reference distance is <= threshold.
 */
int tc_cluster_threshold(tc_ctx *ctx, int kmer_size, int containment, double threshold, uint32_t part, uint32_t nparts,
                         int32_t *labels, uint32_t *n_clusters);
/* host: into[] := labels of the union of the two partial partitions; *n_clusters (may be NULL) = clusters after it */
int tc_labels_merge(int32_t *into, const int32_t *other, uint32_t n, uint32_t *n_clusters);

/* the Boruvka half of tc_mst alone: MST of the candidate edges the last tc_allpairs_pass / _from / _b left on the device
 * (the reference's final Kruskal, src/MST.cpp:1715-1723; `containment` = TC_DIST_* mode matching that pass).  For callers
 * that use the edges in between (tc_dense_ani).  In a multi-GPU group: collective, every rank brings the edges of its part. */
int tc_mst_of_edges(tc_ctx *ctx, int kmer_size, int containment, tc_edge *out, uint64_t *n_out);

/* timings and counts of the last tc_mst / tc_mst_append on the context (bench accounting): out[0] dense-pass kernel ms,
 * [1] Boruvka kernel ms, [2] forest sort ms (CUDA events on tc_stream), [3] candidate edges of this rank, [4] Boruvka
 * rounds, [5] tiles the pass visits (-1: tile flags off), [6] tiles in the lower triangle, [7] 1 = pass on the reordered copy,
 * [8] batches and [9] forest reductions of a bounded-memory build (0: the candidate graph fitted the edge buffer; otherwise
 * the pass ran in batches sized to the buffer -- TCLUST_EDGE_CAP entries or what the device has free -- and Boruvka folded
 * every full buffer into a forest of <= n-1 edges, as the reference does per task, src/MST.cpp:1373-1386) */
int tc_mst_stats(const tc_ctx *ctx, double *out /* 12 doubles */);

/*
 * Multi-GPU group ("N x N tile space sharded over the GPUs of one box, min-reduction per Boruvka round").
 * One context per GPU -- in one process (one host thread per context) or in one process per GPU (torchrun).  Every
 * context exports a handle to its exchange buffer, the host gathers the handles of all ranks in rank order (any
 * plumbing: torch.distributed all_gather, a shared array) and every context joins with the full list.  From then on
 * tc_mst / tc_mst_append on the member contexts are COLLECTIVE calls: every rank must hold the same sketch set (same
 * n <= n_max) and call them in the same order.  Rank r evaluates the work units u with u mod G == r of the dense pass;
 * the per-round min-reductions of Boruvka run inside the kernel over NVLink peer memory (k4_boruvka.cu), with no NCCL
 * call and no host round trip per round; every rank returns the same MST.  Replaces the OpenMP task pool plus final
 * Kruskal of compute_minhash_mst (src/MST.cpp:1373-1386, 1715-1723) across GPUs.  At most 8 ranks.
 */
#define TC_MG_HANDLE_BYTES 128
int tc_mg_export(tc_ctx *ctx, uint32_t n_max, void *handle_out /* TC_MG_HANDLE_BYTES */);
int tc_mg_join(tc_ctx *ctx, int rank, int nranks, const void *handles /* nranks x TC_MG_HANDLE_BYTES, rank order */);
int tc_mg_leave(tc_ctx *ctx);
int tc_mg_rank(const tc_ctx *ctx, int *rank, int *nranks);

/*
 * Incremental MST = append_clust_mst (src/sub_command.cpp:1578-1706): the uploaded set is old genomes [0, start_index)
 * followed by new ones; pre_mst is the saved MST of the old genomes (edge.mst).  Only the pairs (new, earlier) are
 * evaluated on the device; the result is the MST of old-MST + new candidate edges, sorted by (dist, i, j).
 */
int tc_mst_append(tc_ctx *ctx, int kmer_size, int containment, double threshold, uint32_t start_index,
                  const tc_edge *pre_mst, uint64_t n_pre, tc_edge *out, uint64_t *n_out);

/*
 * --dense epilogue of compute_minhash_mst / compute_kssd_mst (src/MST.cpp:1517-1530, 1702-1712) over the candidate
 * pairs of the last tc_allpairs_pass / tc_mst on this context: dense[t*n + v] = number of candidate neighbours of
 * genome v with distance <= t/dense_span (lower_bound bucketing made cumulative over t, the layout of mst.dense,
 * src/MST_IO.cpp:219-233); ani[b], b = (int)((1-dist)*100) clamped to 100 (mst.ani, :235-250).  Bit-exact: bins are
 * decided on exact J = common/denom against host-computed thresholds of the reference's glibc expression.
 */
int tc_dense_ani(tc_ctx *ctx, int kmer_size, int containment, int dense_span, int32_t *dense, uint64_t *ani);

/*
 * Brute-force MST = modifyMST (src/MST.cpp:809-1018): every pair gets MinHash::distance (semantic B with
 * `sketch_size`; containment != 0: containDistance) and the MST of the COMPLETE graph is returned: n-1 edges
 * {i, j<i, dist} sorted by (dist, i, j).  Pairs without a shared hash have distance exactly 1; they are not
 * enumerated but enter as the weight-1 edges that join the remaining components.
 */
int tc_mst_bruteforce(tc_ctx *ctx, int kmer_size, uint32_t sketch_size, int containment,
                      tc_edge *out, uint64_t *n_out);

/*
 * Greedy representative scan (K5).  For queries q in [q0,q1) against representatives
 * rep_ids[0..nrep): common[(q-q0)*nrep + r] and first_hit (index in the query sketch of the
 * first shared hash, UINT32_MAX if none) -- what greedy.cpp:1148-1164 accumulates in
 * cnt[]/touched[].
 */
int tc_greedy_scan(tc_ctx *ctx, uint32_t q0, uint32_t q1, const uint32_t *rep_ids,
                   uint32_t nrep, uint32_t *common, uint32_t *first_hit);
/* full MinHashGreedyClusterWithInvertedIndex with one thread (greedy.cpp:986-1399):
 * rep_of[j] = representative of genome j.  Returns #clusters in *n_clusters. */
int tc_greedy_cluster(tc_ctx *ctx, int kmer_size, int containment, double threshold,
                      int32_t *rep_of, uint32_t *n_clusters);

/* KssdGreedyClusterWithInvertedIndex (src/greedy.cpp:566-899) with one thread, for sketches the caller has already
 * put in the reference's processing order (hash count descending, the reference's std::sort; see the CLI):
 * per-pair common_min, the representative with the largest Jaccard wins, ties -> first in `touched` order. */
int tc_greedy_cluster_kssd(tc_ctx *ctx, int kmer_size, double threshold, int32_t *rep_of, uint32_t *n_clusters);

/* KssdIncrementalCluster (src/greedy.cpp:1736-1880; caller: append_clust_greedy_fast, src/sub_command.cpp:192-276): the
 * uploaded sketches [0, start_index) are already clustered and rep_ids[0..n_reps) are their representatives in creation
 * order (KssdClusterState::representative_ids); every sketch q >= start_index, in order, joins the representative with the
 * smallest Mash distance <= threshold among those passing the two-sided size-ratio filter (2*exp(d*k)-1) and the
 * floor()ed common-hash filter, or becomes the next representative.  rep_of has n entries; only [start_index, n) are
 * written: the sketch id of the chosen representative, or q itself.  *n_clusters = representatives afterwards.
 * Equal distances: the reference walks its candidates in the iteration order of its
 * `phmap::flat_hash_map<int,int> candidate_counts` (reserve(1000)) and keeps the first of the smallest distance; that
 * order is a function of the touched representatives in first-touch order and is reproduced exactly (the device
 * returns every touched representative for these rules).  (Sketches of the reference's single-file KSSD mode are
 * stored there in hash-set order, and its incremental step then merges those unsorted arrays as if they were sorted:
 * its distances are wrong for such folders; this call always works on ascending sets.) */
int tc_greedy_incremental_kssd(tc_ctx *ctx, int kmer_size, double threshold, uint32_t start_index, const uint32_t *rep_ids,
                               uint32_t n_reps, int32_t *rep_of, uint32_t *n_clusters);
/* MinHashIncrementalCluster (src/greedy.cpp:1963-2130; caller: append_clust_greedy with a saved state,
 * src/sub_command.cpp:97-190): as above for MinHash sketches -- no size-ratio filter; containment != 0 uses
 * floor(jmin * min(|A|,|B|)) as the common-hash filter and common / min(|A|,|B|) as the Jaccard. */
int tc_greedy_incremental(tc_ctx *ctx, int kmer_size, int containment, double threshold, uint32_t start_index,
                          const uint32_t *rep_ids, uint32_t n_reps, int32_t *rep_of, uint32_t *n_clusters);

/* The same two rules with the representative scan sharded over several GPUs (the outer loop over the
 * genomes is serial, the scan of one batch against the representatives is not): ctxs[0..nctx) are contexts, one per
 * device, that all hold the SAME uploaded sketch set; representative number c is kept by context c mod nctx, every
 * context scans the batch against its share on its own host thread (the j-parallel loop of greedy.cpp:1136-1336 across
 * devices), the candidates are merged on the host and the serial replay is the single-GPU one -- the result does not
 * depend on nctx.  kssd != 0: the rule of tc_greedy_cluster_kssd (containment ignored).  One caller thread. */
int tc_greedy_cluster_multi(tc_ctx *const *ctxs, int nctx, int kmer_size, int containment, double threshold, int kssd,
                            int32_t *rep_of, uint32_t *n_clusters);

/* brute-force greedyCluster (src/greedy.cpp:285-351) with one thread: genome j joins the nearest representative
 * with MinHash::distance (containment: containDistance) <= threshold, ties -> the earliest-created one; else it
 * becomes a representative.  threshold < 1. */
int tc_greedy_cluster_bruteforce(tc_ctx *ctx, int kmer_size, uint32_t sketch_size, int containment,
                                 double threshold, int32_t *rep_of, uint32_t *n_clusters);

/* ---- host-side exact distance helpers (glibc log; reference expression order) --------- */
double tc_mash_dist(int common, int size0, int size1, int kmer_size);     /* MST.cpp:1489-1503 */
double tc_aaf_dist(int common, int size0, int size1, int kmer_size);      /* MST.cpp:1504-1515 */
double tc_dist_from_key(uint64_t key, int containment, int kmer_size);    /* J bits -> dist, mode TC_DIST_* */
double tc_jaccard_b(int common_b, int size0, int size1, int sketch_size); /* MinHash.cpp:1079-1100 */
double tc_distance_b(int common_b, int size0, int size1, int sketch_size,
                     int kmer_size);                                      /* MinHash.cpp:1125-1153 */
double tc_contain_distance(int common, int size0, int size1, int kmer_size); /* MinHash.cpp:908-979,1102-1123 */
int    tc_greedy_common_min(double threshold, int kmer_size, int size_ref, int size_qry,
                            int containment);                             /* greedy.cpp:1107-1113,1211-1219 */
double tc_greedy_dist(int common, int size_ref, int size_qry, int containment,
                      int kmer_size);                                     /* greedy.cpp:1245-1274 */
int    tc_radio(double threshold, int kmer_size);                         /* MST.cpp:26-37,1292 */

#ifdef __cplusplus
}
#endif
#endif
