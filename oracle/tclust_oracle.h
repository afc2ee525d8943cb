/*
 * tclust_oracle.h -- CPU restatement ("oracle") of the RabbitTClust hot paths.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 * The product path (rabbittclust_b200/, include/tclust.h) never links or calls it.
 *
 * Parity pinning: every function below is checked (tests/test_oracle_*.py) against
 * outputs of the reference itself, compiled from /root/reference by oracle/Makefile
 * into oracle/_ref/ (clust-mst, clust-greedy, ref_harness), and against the golden
 * vectors those binaries produced (tests/golden/, generator tools/make_golden.py).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef TCLUST_ORACLE_H
#define TCLUST_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t pre, suf; double dist; } oc_edge;   /* src/MST.h:17-21 */

/* RabbitSketch/src/MurmurHash3.cpp:259-336; returns out[0] (what MinHash keeps,
 * MinHash.cpp:350-377). */
uint64_t oc_murmur3_x64_128_h1(const uint8_t *key, int len, uint32_t seed);

/* RabbitSketch/src/MinHash.cpp:1185-1195 */
void oc_revcomp(const uint8_t *src, uint8_t *dst, size_t n);

/* MinHash::update + MinHashHeap + heapToList (MinHash.cpp:465-747, 287-348, 756-857):
 * one genome = nrec records seq[rec_off[r] .. rec_off[r+1]); k-mers never span
 * records; upper-case; canonical strand by rule 0 = memcmp (AVX2 build / Mash),
 * rule 1 = AVX-512 build's min512 lane rule for the 16-aligned body (test-only).
 * Output: the `s` smallest distinct hashes ascending; returns their count. */
size_t oc_minhash_sketch(const uint8_t *seq, const uint64_t *rec_off, int nrec,
                         int k, size_t s, uint32_t seed, int rule, uint64_t *out);

/* KSSD list-mode sketch, src/SketchInfo.cpp:994-1252 (hot loop 1126-1165);
 * shuffled_dim from oc_kssd_shuffled_dim. Returns count; out is sorted ascending,
 * values are u32-ranged unless use64 (half_k - drlevel > 8). */
void   oc_kssd_shuffled_dim(int half_subk, int32_t *out /* 1<<(4*half_subk) */); /* SketchInfo.cpp:60-102 */
size_t oc_kssd_sketch(const uint8_t *seq, const uint64_t *rec_off, int nrec,
                      int kmer_size, int drlevel, const int32_t *shuffled_dim,
                      uint64_t *out, size_t out_cap);

/* the single-file (-i) consumer, one record: consumer_fasta_task_with_kssd (SketchInfo.cpp:274-433); equals
 * oc_kssd_sketch for an even kmer_size, keeps one stale base per reset and the un-rounded shift for an odd one */
size_t oc_kssd_sketch_seq(const uint8_t *seq, uint64_t len, int kmer_size, int drlevel, const int32_t *shuffled_dim,
                          uint64_t *out, size_t out_cap);

/* semantic A: |A n B| over the whole sorted sketches (src/MST.cpp:1409-1435 counts,
 * scalar merge form src/greedy.cpp:125-147). first_hit (may be NULL) = index in `a`
 * of the first shared hash, or UINT32_MAX. */
uint32_t oc_intersect_u64(const uint64_t *a, size_t na, const uint64_t *b, size_t nb,
                          uint32_t *first_hit);

/* semantic B: Mash union-limited Jaccard, MinHash.cpp:982-1101 (scalar form). */
double oc_jaccard_B(const uint64_t *a, size_t na, const uint64_t *b, size_t nb,
                    uint64_t sketch_size);
double oc_distance_B(double jaccard, int k);                      /* MinHash.cpp:1125-1153 */

int    oc_radio(double threshold, int kmer_size);                  /* MST.cpp:26-37,1292 */
double oc_mash_dist_A(int common, int size0, int size1, int k);    /* MST.cpp:1489-1503 */
double oc_aaf_dist_mst(int common, int size0, int size1, int k);   /* MST.cpp:1504-1515 */
double oc_greedy_dist(int common, int size_ref, int size_qry, int containment, int k); /* greedy.cpp:1245-1274 */
int    oc_greedy_common_min(double threshold, int k, int size_ref, int size_qry,
                            int containment);                      /* greedy.cpp:1107-1113,1211-1219 */

/* compute_minhash_mst (MST.cpp:1290-1737): edges {i, j<i, dist} for pairs with
 * common>=1 passing the size-ratio filter; sort by (dist,i,j); Kruskal
 * (MST.cpp:59-75).  Returns number of forest edges written to out (cap N-1). */
size_t oc_mst(const uint64_t *hashes, const uint64_t *off, int N, int k,
              int containment, double threshold, oc_edge *out);

/* all-pairs common matrix (row-major N x N, diagonal = size) for tests. */
void oc_common_matrix(const uint64_t *hashes, const uint64_t *off, int N, uint32_t *out);

/* generateForest + generateClusterWithBfs (MST.cpp:77-85,109-142): labels[v] =
 * cluster index in reference order (clusters ordered by smallest vertex); order[] =
 * vertices in output order (cluster by cluster, BFS order). Returns #clusters. */
int oc_clusters_from_mst(const oc_edge *mst, size_t n_edges, int N, double threshold,
                         int32_t *labels, int32_t *order);

/* MinHashGreedyClusterWithInvertedIndex with -t 1 (greedy.cpp:986-1399), including
 * the touched-order tie rule. rep_of[j] = representative genome id
 * (itself if j is a representative). Returns #clusters. */
int oc_greedy_inverted(const uint64_t *hashes, const uint64_t *off, int N, int k,
                       int containment, double threshold, int32_t *rep_of);

/* KssdGreedyClusterWithInvertedIndex with -t 1 (greedy.cpp:566-899) on sketches in its processing order (size
 * descending): per-pair ceil()ed common filter, largest Jaccard wins, first in touched order among equals. */
int oc_greedy_kssd(const uint64_t *hashes, const uint64_t *off, int N, int k, double threshold, int32_t *rep_of);

/* iteration order of `phmap::flat_hash_map<int,int> m; m.reserve(1000); m[key]++` for distinct keys in insertion order
 * (the reference's candidate table, greedy.cpp:1771-1803; src/phmap.h) */
void oc_phmap_iteration_order(const int32_t *keys, int n, int32_t *out);

/* KssdIncrementalCluster with -t 1 (greedy.cpp:1736-1880): sketches [0,start) clustered with representatives
 * seed[0..n_seed) (creation order); rep_of[q], q >= start: chosen representative or q.  Equal distances: the
 * candidate that comes first in the iteration order of the reference's candidate table. Returns #representatives. */
int oc_greedy_incremental_kssd(const uint64_t *hashes, const uint64_t *off, int N, int k, double threshold,
                               int start, const int32_t *seed, int n_seed, int32_t *rep_of);

/* MinHashIncrementalCluster with -t 1 (greedy.cpp:1963-2130): no size-ratio filter, containment forms */
int oc_greedy_incremental_minhash(const uint64_t *hashes, const uint64_t *off, int N, int k, int containment,
                                  double threshold, int start, const int32_t *seed, int n_seed, int32_t *rep_of);

/* ---- semantic B (the brute-force fallbacks) ---------------------------------------- */
/* integer form of MinHash::jaccard (MinHash.cpp:982-1101): common among the first
 * sketch_size union elements; *denom_out = completed/capped denominator. */
uint32_t oc_common_B(const uint64_t *a, size_t na, const uint64_t *b, size_t nb,
                     uint64_t sketch_size, uint32_t *denom_out);
/* MinHash::containDistance (MinHash.cpp:1102-1123, containJaccard :908-979) */
double oc_contain_distance(const uint64_t *a, size_t na, const uint64_t *b, size_t nb, int k);
/* modifyMST (src/MST.cpp:809-1018) + Kruskal: brute-force all pairs with MinHash::distance
 * (containment: containDistance); N-1 edges {i, j>i, dist} in Kruskal order. */
size_t oc_mst_bruteforce(const uint64_t *hashes, const uint64_t *off, int N, int k,
                         uint64_t sketch_size, int containment, oc_edge *out);
/* greedyCluster (src/greedy.cpp:285-351) with one thread. rep_of[j] = centre of genome j. */
int oc_greedy_bruteforce(const uint64_t *hashes, const uint64_t *off, int N, int k,
                         uint64_t sketch_size, int containment, double threshold, int32_t *rep_of);

/* ---- --dense epilogue and noise removal ------------------------------------------------ */
void oc_dense_ani(const uint64_t *hashes, const uint64_t *off, int N, int k, int containment,
                  double threshold, int span, int32_t *dense /* span x N */, uint64_t *ani /* 101 */);
int oc_remove_noise_clusters(const oc_edge *mst, size_t n_edges, int N, double threshold, const int32_t *dense,
                             int span, int32_t *labels, int32_t *order);

#ifdef __cplusplus
}
#endif
#endif
