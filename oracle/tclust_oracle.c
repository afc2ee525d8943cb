/*
 * tclust_oracle.c -- CPU restatement of the RabbitTClust hot paths (plain C).
 * TEST INFRASTRUCTURE ONLY -- see tclust_oracle.h.  Not linked by the product.
 * Citations are relative to /root/reference.
 */
#include "tclust_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ Murmur3 */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t k) {           /* MurmurHash3.cpp fmix64 */
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33; return k;
}

/* RabbitSketch/src/MurmurHash3.cpp:259-336 (public-domain MurmurHash3_x64_128). */
uint64_t oc_murmur3_x64_128_h1(const uint8_t *data, int len, uint32_t seed)
{
    const int nblocks = len / 16;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (int i = 0; i < nblocks; i++) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * i, 8);
        memcpy(&k2, data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    int rem = len & 15;
    for (int i = rem - 1; i >= 8; i--) k2 |= (uint64_t)tail[i] << (8 * (i - 8));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int i = (rem > 8 ? 7 : rem - 1); i >= 0; i--) k1 |= (uint64_t)tail[i] << (8 * i);
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; /* h2 += h1; */
    return h1;
}

/* RabbitSketch/src/MinHash.cpp:1185-1195 */
void oc_revcomp(const uint8_t *src, uint8_t *dst, size_t n)
{
    static const uint8_t table[4] = {'T', 'G', 'A', 'C'};
    for (size_t i = 0; i < n; i++) dst[n - 1 - i] = table[(src[i] >> 1) & 3];
}

/* ------------------------------------------------------- bottom-s selection */
static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
typedef struct { uint64_t *buf; size_t n, cap, s; uint64_t tau; } bottom_s;

/* sort + unique + truncate to s; tau = s-th smallest once s distinct values exist.
 * Equivalent to MinHashHeap::tryInsert's "s smallest distinct" (MinHash.cpp:287-348). */
static void bs_compact(bottom_s *b)
{
    qsort(b->buf, b->n, 8, cmp_u64);
    size_t m = 0;
    for (size_t i = 0; i < b->n; i++)
        if (m == 0 || b->buf[i] != b->buf[m - 1]) b->buf[m++] = b->buf[i];
    if (m > b->s) m = b->s;
    b->n = m;
    if (m == b->s && m > 0) b->tau = b->buf[m - 1];
}
static inline void bs_push(bottom_s *b, uint64_t h)
{
    if (h > b->tau) return;           /* cannot be among the s smallest distinct */
    if (b->n == b->cap) bs_compact(b);
    b->buf[b->n++] = h;
}

/* AVX-512 build's strand choice, MinHash.cpp:396-402 (min512) + 555-571: lanes of
 * 8 bytes of the zero-padded k-mer, highest lane first, signed 64-bit compare. */
static int fwd_is_min512(const uint8_t *f, const uint8_t *r, int k)
{
    uint8_t a[64], b[64];
    memset(a, 0, 64); memset(b, 0, 64);
    memcpy(a, f, k); memcpy(b, r, k);
    for (int lane = 7; lane >= 0; lane--) {
        int64_t x, y;
        memcpy(&x, a + 8 * lane, 8); memcpy(&y, b + 8 * lane, 8);
        if (x < y) return 1;
        if (x > y) return 0;
    }
    return 0;
}

/* MinHash::update (MinHash.cpp:465-747) over every record, then heapToList
 * (756-810). */
size_t oc_minhash_sketch(const uint8_t *seq, const uint64_t *rec_off, int nrec,
                         int k, size_t s, uint32_t seed, int rule, uint64_t *out)
{
    bottom_s b;
    b.s = s; b.cap = 4 * s + 1024; b.n = 0; b.tau = UINT64_MAX;
    b.buf = (uint64_t *)malloc(b.cap * 8);
    for (int r = 0; r < nrec; r++) {
        size_t len = (size_t)(rec_off[r + 1] - rec_off[r]);
        if (len < (size_t)k) continue;
        uint8_t *fw = (uint8_t *)malloc(len), *rc = (uint8_t *)malloc(len);
        memcpy(fw, seq + rec_off[r], len);
        for (size_t i = 0; i < len; i++)                 /* MinHash.cpp:472-479 */
            if (fw[i] > 96 && fw[i] < 123) fw[i] -= 32;
        oc_revcomp(fw, rc, len);
        size_t n = len - k + 1, body16 = (n / 16) * 16;
        for (size_t i = 0; i < n; i++) {
            const uint8_t *f = fw + i, *v = rc + len - i - k;
            int use_fwd;
            if (rule == 1 && i < body16) use_fwd = fwd_is_min512(f, v, k);
            else use_fwd = memcmp(f, v, k) <= 0;          /* MinHash.cpp:667-670,696 */
            bs_push(&b, oc_murmur3_x64_128_h1(use_fwd ? f : v, k, seed));
        }
        free(fw); free(rc);
    }
    bs_compact(&b);
    memcpy(out, b.buf, b.n * 8);
    size_t n = b.n;
    free(b.buf);
    return n;
}

/* --------------------------------------------------------------------- KSSD */
static void oc_shuffle(int32_t *arr, int length, unsigned seed)  /* SketchInfo.cpp:60-78 */
{
    srand(seed);
    for (int i = length - 1; i > 0; i--) {
        int j = rand() % (i + 1);
        int32_t t = arr[i]; arr[i] = arr[j]; arr[j] = t;
    }
}
void oc_kssd_shuffled_dim(int half_subk, int32_t *out)          /* SketchInfo.cpp:80-102 */
{
    int n = 1 << (4 * half_subk);
    for (int i = 0; i < n; i++) out[i] = i;
    oc_shuffle(out, n, 23);
    oc_shuffle(out, n, 348842630u);
}

size_t oc_kssd_sketch(const uint8_t *seq, const uint64_t *rec_off, int nrec,
                      int kmer_size, int drlevel, const int32_t *shuffled_dim,
                      uint64_t *out, size_t out_cap)
{
    /* constants: SketchInfo.cpp:1019-1048 */
    int half_k = (kmer_size + 1) / 2;
    int kmer = half_k * 2;
    int use64 = half_k - drlevel > 8;
    int half_subk = 6 - drlevel >= 2 ? 6 : drlevel + 2;
    int dim_end = 1 << (4 * (half_subk - drlevel));
    int comp_bittl = 64 - 4 * half_k;
    int half_outctx_len = half_k - half_subk;
    int rev_add_move = 4 * half_k - 2;
    uint64_t tupmask = UINT64_MAX >> comp_bittl;
    uint64_t domask = (tupmask >> (4 * half_outctx_len)) << (2 * half_outctx_len);
    uint64_t undomask = (tupmask ^ domask) & tupmask;
    uint64_t undomask1 = undomask & (tupmask >> ((half_k + half_subk) * 2));
    uint64_t undomask0 = undomask ^ undomask1;
    size_t n = 0, cap = 1 << 16;
    uint64_t *buf = (uint64_t *)malloc(cap * 8);
    for (int r = 0; r < nrec; r++) {
        uint64_t tuple = 0, rvs = 0;
        int base = 1;
        for (uint64_t p = rec_off[r]; p < rec_off[r + 1]; p++) {   /* :1126-1165 */
            int bn;
            switch (seq[p]) {
                case 'A': case 'a': bn = 0; break;
                case 'C': case 'c': bn = 1; break;
                case 'G': case 'g': bn = 2; break;
                case 'T': case 't': bn = 3; break;
                default: bn = -1;
            }
            if (bn < 0) { base = 1; tuple = 0; rvs = 0; continue; }
            tuple = ((tuple << 2) | (uint64_t)bn) & tupmask;
            rvs = (rvs >> 2) + (((uint64_t)bn ^ 3ULL) << rev_add_move);
            base++;
            if (base > kmer) {
                uint64_t uni = tuple < rvs ? tuple : rvs;
                uint32_t dim_id = (uint32_t)((uni & domask) >> (half_outctx_len * 2));
                int pf = shuffled_dim[dim_id];
                if (pf >= dim_end) continue;
                uint64_t dr = (((uni & undomask0) |
                                ((uni & undomask1) << (kmer * 2 - half_outctx_len * 4)))
                               >> (drlevel * 4)) | (uint64_t)pf;
                if (!use64) dr &= 0xffffffffULL;
                if (n == cap) { cap *= 2; buf = (uint64_t *)realloc(buf, cap * 8); }
                buf[n++] = dr;
            }
        }
    }
    qsort(buf, n, 8, cmp_u64);                                     /* :1180-1193 */
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
        if (m == 0 || buf[i] != buf[m - 1]) buf[m++] = buf[i];
    if (m > out_cap) m = out_cap;
    memcpy(out, buf, m * 8);
    free(buf);
    return m;
}

/* The per-record consumer of the single-file mode (consumer_fasta_task_with_kssd, SketchInfo.cpp:274-433).  It differs
 * from the list-mode loop above in three ways, all of which matter only for an ODD kmer_size: kmerSize is NOT rounded up
 * to 2*half_k (:296 vs :1019-1020), so a window is used once kmer_size valid bases have been seen (`base > kmerSize`,
 * :372) while the tuple still holds 2*half_k bases; a non-ACGT byte restarts the count but does NOT clear tuple /
 * rvs_tuple (:368-371), so the first window after it carries one base from before the gap (or the zero initial state at
 * the start of the record); and the outer-context shift uses the un-rounded kmerSize (:386).  One sketch per record;
 * the reference stores the set in hash-set iteration order, returned here ascending. */
size_t oc_kssd_sketch_seq(const uint8_t *seq, uint64_t len, int kmer_size, int drlevel, const int32_t *shuffled_dim,
                          uint64_t *out, size_t out_cap)
{
    int half_k = (kmer_size + 1) / 2;
    int use64 = half_k - drlevel > 8;
    int half_subk = 6 - drlevel >= 2 ? 6 : drlevel + 2;
    int dim_end = 1 << (4 * (half_subk - drlevel));
    int comp_bittl = 64 - 4 * half_k;
    int half_outctx_len = half_k - half_subk;
    int rev_add_move = 4 * half_k - 2;
    uint64_t tupmask = UINT64_MAX >> comp_bittl;
    uint64_t domask = (tupmask >> (4 * half_outctx_len)) << (2 * half_outctx_len);
    uint64_t undomask = (tupmask ^ domask) & tupmask;
    uint64_t undomask1 = undomask & (tupmask >> ((half_k + half_subk) * 2));
    uint64_t undomask0 = undomask ^ undomask1;
    size_t n = 0, cap = 1 << 12;
    uint64_t *buf = (uint64_t *)malloc(cap * 8);
    uint64_t tuple = 0, rvs = 0;
    int base = 1;
    for (uint64_t p = 0; p < len; p++) {                           /* :358-394 */
        int bn;
        switch (seq[p]) {
            case 'A': case 'a': bn = 0; break;
            case 'C': case 'c': bn = 1; break;
            case 'G': case 'g': bn = 2; break;
            case 'T': case 't': bn = 3; break;
            default: bn = -1;
        }
        if (bn >= 0) {
            tuple = ((tuple << 2) | (uint64_t)bn) & tupmask;
            rvs = (rvs >> 2) + (((uint64_t)bn ^ 3ULL) << rev_add_move);
            base++;
        } else {
            base = 1;
        }
        if (base > kmer_size) {            /* also true at a non-ACGT byte?  no: base was just set to 1 */
            uint64_t uni = tuple < rvs ? tuple : rvs;
            uint32_t dim_id = (uint32_t)((uni & domask) >> (half_outctx_len * 2));
            int pf = shuffled_dim[dim_id];
            if (pf >= dim_end) continue;
            uint64_t dr = (((uni & undomask0) | ((uni & undomask1) << (kmer_size * 2 - half_outctx_len * 4)))
                           >> (drlevel * 4)) | (uint64_t)pf;
            if (!use64) dr &= 0xffffffffULL;                       /* stored in a uint32 set */
            if (n == cap) { cap *= 2; buf = (uint64_t *)realloc(buf, cap * 8); }
            buf[n++] = dr;
        }
    }
    qsort(buf, n, 8, cmp_u64);
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
        if (m == 0 || buf[i] != buf[m - 1]) buf[m++] = buf[i];
    if (m > out_cap) m = out_cap;
    memcpy(out, buf, m * 8);
    free(buf);
    return m;
}

/* --------------------------------------------------------------- distances */
uint32_t oc_intersect_u64(const uint64_t *a, size_t na, const uint64_t *b, size_t nb,
                          uint32_t *first_hit)
{
    size_t i = 0, j = 0; uint32_t c = 0, fh = UINT32_MAX;
    while (i < na && j < nb) {
        if (a[i] < b[j]) i++;
        else if (a[i] > b[j]) j++;
        else { if (!c) fh = (uint32_t)i; c++; i++; j++; }
    }
    if (first_hit) *first_hit = fh;
    return c;
}

double oc_jaccard_B(const uint64_t *a, size_t na, const uint64_t *b, size_t nb,
                    uint64_t sketch_size)                 /* MinHash.cpp:1058-1100 */
{
    uint64_t i = 0, j = 0, common = 0, denom = 0;
    while (denom < sketch_size && i < na && j < nb) {
        if (a[i] < b[j]) i++;
        else if (b[j] < a[i]) j++;
        else { i++; j++; common++; }
        denom++;
    }
    if (denom < sketch_size) {
        if (i < na) denom += na - i;
        if (j < nb) denom += nb - j;
        if (denom > sketch_size) denom = sketch_size;
    }
    return (double)common / (double)denom;
}

double oc_distance_B(double jaccard, int k)               /* MinHash.cpp:1125-1153 */
{
    if (jaccard == 1.0) return 0;
    double d = -log(2 * jaccard / (1. + jaccard)) / k;
    if (d > 1) d = 1;
    return d;
}

int oc_radio(double threshold, int kmer_size)             /* MST.cpp:26-37,1292 */
{
    return (int)(2.0 * exp(threshold * (kmer_size - 1)) - 1.0);
}

double oc_mash_dist_A(int common, int size0, int size1, int k)   /* MST.cpp:1489-1503 */
{
    const double inv_k = 1.0 / k;
    int denom = size0 + size1 - common;
    double jaccard = denom == 0 ? 0.0 : (double)common / denom;
    if (jaccard == 1.0) return 0.0;
    if (jaccard == 0.0) return 1.0;
    double ratio = (2.0 * jaccard) / (1.0 + jaccard);
    return -inv_k * log(ratio);
}

double oc_aaf_dist_mst(int common, int size0, int size1, int k)  /* MST.cpp:1504-1515 */
{
    const double inv_k = 1.0 / k;
    int denom = size0 < size1 ? size0 : size1;
    double c = denom == 0 ? 0.0 : (double)common / denom;
    if (c == 1.0) return 0.0;
    if (c == 0.0) return 1.0;
    return -inv_k * log(c);
}

double oc_greedy_dist(int common, int size_ref, int size_qry, int containment, int k)
{                                                          /* greedy.cpp:1245-1274 */
    double dist;
    if (containment) {
        int mn = size_ref < size_qry ? size_ref : size_qry;
        if (mn == 0) return 1.0;
        double j = (double)common / mn;
        if (j >= 1.0) dist = 0.0;
        else if (j <= 0.0) dist = 1.0;
        else { dist = -log(2.0 * j / (1.0 + j)) / k; if (dist > 1.0) dist = 1.0; }
    } else {
        int denom = size_ref + size_qry - common;
        if (denom == 0) return 0.0;
        double j = (double)common / denom;
        if (j >= 1.0) dist = 0.0;
        else if (j <= 0.0) dist = 1.0;
        else { dist = -log(2.0 * j / (1.0 + j)) / k; if (dist > 1.0) dist = 1.0; }
    }
    return dist;
}

int oc_greedy_common_min(double threshold, int k, int size_ref, int size_qry, int containment)
{                                                          /* greedy.cpp:1107-1113,1211-1219 */
    double x = exp(-threshold * k);
    double jmin = x / (2.0 - x);
    if (containment) {
        int mn = size_ref < size_qry ? size_ref : size_qry;
        return (int)ceil(jmin * mn);
    }
    return (int)ceil(jmin * (size_ref + size_qry) / (1.0 + jmin));
}

/* ---------------------------------------------------------------------- MST */
static int uf_find(int *p, int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }

static int cmp_edge(const void *a, const void *b)
{
    const oc_edge *x = (const oc_edge *)a, *y = (const oc_edge *)b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    if (x->pre != y->pre) return x->pre < y->pre ? -1 : 1;
    return x->suf < y->suf ? -1 : x->suf > y->suf;
}

size_t oc_mst(const uint64_t *hashes, const uint64_t *off, int N, int k,
              int containment, double threshold, oc_edge *out)
{
    int radio = oc_radio(threshold, k);
    size_t cap = 1 << 16, n = 0;
    oc_edge *g = (oc_edge *)malloc(cap * sizeof(oc_edge));
    for (int i = 0; i < N; i++) {
        int size0 = (int)(off[i + 1] - off[i]);
        if (size0 == 0) continue;                              /* MST.cpp:1404 */
        for (int j = 0; j < i; j++) {
            int size1 = (int)(off[j + 1] - off[j]);
            if (size1 == 0) continue;
            int common = (int)oc_intersect_u64(hashes + off[i], size0, hashes + off[j], size1, NULL);
            if (common == 0) continue;        /* only index hits are candidates, :1409-1435 */
            int mn = size0 < size1 ? size0 : size1, mx = size0 > size1 ? size0 : size1;
            if (mx > radio * mn) continue;                     /* :1481-1487 */
            double d = containment ? oc_aaf_dist_mst(common, size0, size1, k)
                                   : oc_mash_dist_A(common, size0, size1, k);
            if (n == cap) { cap *= 2; g = (oc_edge *)realloc(g, cap * sizeof(oc_edge)); }
            g[n].pre = i; g[n].suf = j; g[n].dist = d; n++;
        }
    }
    qsort(g, n, sizeof(oc_edge), cmp_edge);                    /* :1721 */
    int *p = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    for (int i = 0; i < N; i++) p[i] = i;
    size_t m = 0;
    for (size_t e = 0; e < n; e++) {                           /* kruskalAlgorithm :59-75 */
        int a = uf_find(p, g[e].pre), b = uf_find(p, g[e].suf);
        if (a != b) { p[a] = b; out[m++] = g[e]; }
    }
    free(p); free(g);
    return m;
}

void oc_common_matrix(const uint64_t *hashes, const uint64_t *off, int N, uint32_t *out)
{
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            out[(size_t)i * N + j] = oc_intersect_u64(hashes + off[i], off[i + 1] - off[i],
                                                      hashes + off[j], off[j + 1] - off[j], NULL);
}

int oc_clusters_from_mst(const oc_edge *mst, size_t n_edges, int N, double threshold,
                         int32_t *labels, int32_t *order)
{
    /* generateForest (MST.cpp:77-85) + generateClusterWithBfs (:109-142) */
    int *deg = (int *)calloc(N + 1, sizeof(int));
    for (size_t e = 0; e < n_edges; e++)
        if (mst[e].dist <= threshold) { deg[mst[e].pre]++; deg[mst[e].suf]++; }
    int *start = (int *)malloc(sizeof(int) * (N + 1));
    start[0] = 0;
    for (int i = 0; i < N; i++) start[i + 1] = start[i] + deg[i];
    int *adj = (int *)malloc(sizeof(int) * (start[N] + 1));
    int *fill = (int *)calloc(N + 1, sizeof(int));
    for (size_t e = 0; e < n_edges; e++)
        if (mst[e].dist <= threshold) {
            int u = mst[e].pre, v = mst[e].suf;
            adj[start[u] + fill[u]++] = v;
            adj[start[v] + fill[v]++] = u;
        }
    char *vis = (char *)calloc(N + 1, 1);
    int *q = (int *)malloc(sizeof(int) * (N + 1));
    int nclust = 0, pos = 0;
    for (int i = 0; i < N; i++) {
        if (vis[i]) continue;
        vis[i] = 1;
        int qh = 0, qt = 0;
        q[qt++] = i; labels[i] = nclust; order[pos++] = i;
        while (qh < qt) {
            int u = q[qh++];
            for (int t = start[u]; t < start[u + 1]; t++) {
                int v = adj[t];
                if (vis[v]) continue;
                vis[v] = 1; q[qt++] = v; labels[v] = nclust; order[pos++] = v;
            }
        }
        nclust++;
    }
    free(deg); free(start); free(adj); free(fill); free(vis); free(q);
    return nclust;
}

/* ------------------------------------------------------------------- greedy */
typedef struct { uint32_t first_hit; int order; int rep; int common; } touched_t;
static int cmp_touched(const void *a, const void *b)
{
    const touched_t *x = (const touched_t *)a, *y = (const touched_t *)b;
    if (x->first_hit != y->first_hit) return x->first_hit < y->first_hit ? -1 : 1;
    return x->order < y->order ? -1 : x->order > y->order;
}

int oc_greedy_inverted(const uint64_t *hashes, const uint64_t *off, int N, int k,
                       int containment, double threshold, int32_t *rep_of)
{
    if (N == 0) return 0;
    int *reps = (int *)malloc(sizeof(int) * N), nrep = 0;
    touched_t *tc = (touched_t *)malloc(sizeof(touched_t) * N);
    /* fixed-size fast path detection over the first 100 sketches, greedy.cpp:1092-1103 */
    int fixed_size = (int)(off[1] - off[0]);
    int all_fixed = !containment;
    for (int i = 1; i < (N < 100 ? N : 100); i++)
        if (containment || (int)(off[i + 1] - off[i]) != fixed_size) { all_fixed = 0; break; }
    int fixed_common_min = 0;
    if (all_fixed) fixed_common_min = oc_greedy_common_min(threshold, k, fixed_size, fixed_size, 0);
    reps[nrep++] = 0; rep_of[0] = 0;
    for (int j = 1; j < N; j++) {
        const uint64_t *q = hashes + off[j];
        int size_ref = (int)(off[j + 1] - off[j]);
        int nt = 0;
        for (int r = 0; r < nrep; r++) {                       /* index probe :1148-1164 */
            int id = reps[r];
            uint32_t fh;
            int c = (int)oc_intersect_u64(q, size_ref, hashes + off[id], off[id + 1] - off[id], &fh);
            if (c > 0) { tc[nt].first_hit = fh; tc[nt].order = r; tc[nt].rep = id; tc[nt].common = c; nt++; }
        }
        qsort(tc, nt, sizeof(touched_t), cmp_touched);         /* touched order, 8a-G */
        int best_common = -1, best_rep = -1;
        double best_dist = 1.7976931348623157e308;
        for (int t = 0; t < nt; t++) {                         /* :1198-1285 with one thread */
            int id = tc[t].rep, common = tc[t].common;
            int size_qry = (int)(off[id + 1] - off[id]);
            int cmin = all_fixed ? fixed_common_min
                                 : oc_greedy_common_min(threshold, k, size_ref, size_qry, containment);
            if (common < cmin) continue;
            if (all_fixed) {
                if (common > best_common) { best_common = common; best_rep = id; }
            } else {
                double d = oc_greedy_dist(common, size_ref, size_qry, containment, k);
                if (d <= threshold && d < best_dist) { best_dist = d; best_common = common; best_rep = id; }
            }
        }
        if (best_rep != -1) rep_of[j] = best_rep;
        else { reps[nrep++] = j; rep_of[j] = j; }
    }
    free(reps); free(tc);
    return nrep;
}

/* KssdGreedyClusterWithInvertedIndex with one thread (greedy.cpp:566-899), for sketches already in its processing order
 * (hash count descending, its std::sort :592-596): sketch j probes the index of the representatives (:707-735), a touched
 * representative passes when common >= ceil(jmin * (|A| + |B|) / (1 + jmin)) (:771), the one of largest Jaccard wins, the
 * first in `touched` order among equals (strict >, :790-793; `touched` = order of first contact: query hashes ascending,
 * postings in representative order).  The periodic pruning of too-large representatives (:690-704) only removes
 * representatives that can no longer pass the filter (|rep| > |query| / (0.8 jmin) implies common_min > |query|).
 * rep_of[j] = representative sketch of j (itself for a representative).  Returns #clusters. */
int oc_greedy_kssd(const uint64_t *hashes, const uint64_t *off, int N, int k, double threshold, int32_t *rep_of)
{
    if (N == 0) return 0;
    int *reps = (int *)malloc(sizeof(int) * N), nrep = 0;
    touched_t *tc = (touched_t *)malloc(sizeof(touched_t) * N);
    const double x = exp(-threshold * k);                    /* :650-652 */
    const double jaccard_min = x / (2.0 - x);
    reps[nrep++] = 0; rep_of[0] = 0;
    for (int j = 1; j < N; j++) {
        const uint64_t *q = hashes + off[j];
        const int sizeRef = (int)(off[j + 1] - off[j]);
        int nt = 0;
        for (int r = 0; r < nrep; r++) {
            const int id = reps[r];
            uint32_t fh;
            const int c = (int)oc_intersect_u64(q, sizeRef, hashes + off[id], off[id + 1] - off[id], &fh);
            if (c > 0) { tc[nt].first_hit = fh; tc[nt].order = r; tc[nt].rep = id; tc[nt].common = c; nt++; }
        }
        qsort(tc, nt, sizeof(touched_t), cmp_touched);
        double best_jaccard = -1.0;
        int best_rep = -1;
        for (int t = 0; t < nt; t++) {
            const int id = tc[t].rep, common = tc[t].common;
            const int sizeQry = (int)(off[id + 1] - off[id]);
            const int common_min = (int)ceil(jaccard_min * (sizeRef + sizeQry) / (1.0 + jaccard_min));
            if (common < common_min) continue;
            const int denom = sizeRef + sizeQry - common;
            const double jaccard = denom == 0 ? 1.0 : (double)common / denom;
            if (jaccard > best_jaccard) { best_jaccard = jaccard; best_rep = id; }
        }
        if (best_rep != -1) rep_of[j] = best_rep;
        else { reps[nrep++] = j; rep_of[j] = j; }
    }
    free(reps); free(tc);
    return nrep;
}

/* ---- iteration order of the reference's candidate table ----------------------------------------------------------
 * Both incremental rules collect the touched representatives in `phmap::flat_hash_map<int,int> candidate_counts;
 * candidate_counts.reserve(1000)` (greedy.cpp:1771-1772, 2014-2015), copy the table into a vector in ITERATION order and
 * walk that vector with a strict `dist < best_dist` (:1805-1848, 2038-2094): among equal distances the candidate that comes
 * first in the table wins.  The table is the vendored parallel-hashmap (src/phmap.h, deterministic build: H1 = hash >> 7,
 * no per-table seed): hash = hi64 + lo64 of key * 0xde5fb9d2630458e9 (phmap_utils.h:67-77), 16-wide groups, triangular
 * probing (phmap.h:145-167), first empty slot of the first group that has one (:2196-2207), capacity 2047 after
 * reserve(1000) (NormalizeCapacity(1000 + 999/7), :573-609, :1698-1714), growth to 2*cap+1 when cap - cap/8 entries are in,
 * re-inserting in slot order (:2054-2079, :2144-2154, :2252-2269).  Checked against the real container through
 * `ref_harness phorder` (tests/test_oracle_golden.py). */
typedef struct { int cap, size, growth; unsigned char *full; int *key; } phtab;
static uint64_t ph_mix(uint64_t a)
{
    unsigned __int128 p = (unsigned __int128)a * 0xde5fb9d2630458e9ULL;
    return (uint64_t)(p >> 64) + (uint64_t)p;
}
static int ph_first_non_full(const phtab *t, uint64_t h)
{
    const uint64_t mask = (uint64_t)t->cap;
    uint64_t off = (h >> 7) & mask, idx = 0;
    for (;;) {
        for (int i = 0; i < 16; i++) {
            const uint64_t p = off + i;            /* control bytes: [0,cap) slots, cap = sentinel, then clones of the first 16 */
            if (p == mask) continue;
            const uint64_t q = p < mask ? p : p - mask - 1;
            if (!t->full[q]) return (int)((off + i) & mask);
        }
        idx += 16; off = (off + idx) & mask;
    }
}
static void ph_resize(phtab *t, int newcap)
{
    unsigned char *of = t->full; int *ok = t->key; const int oc = t->cap;
    t->cap = newcap;
    t->full = (unsigned char *)calloc((size_t)newcap, 1);
    t->key = (int *)malloc(sizeof(int) * (size_t)newcap);
    t->growth = newcap - newcap / 8 - t->size;
    for (int i = 0; i < oc; i++)
        if (of[i]) { const int s = ph_first_non_full(t, ph_mix((uint64_t)ok[i])); t->full[s] = 1; t->key[s] = ok[i]; }
    free(of); free(ok);
}
/* keys: distinct ints in first-insertion order; out: the same keys in iteration order */
void oc_phmap_iteration_order(const int32_t *keys, int n, int32_t *out)
{
    phtab t = {0, 0, 0, NULL, NULL};
    ph_resize(&t, 2047);
    for (int i = 0; i < n; i++) {
        int s = ph_first_non_full(&t, ph_mix((uint64_t)keys[i]));
        if (t.growth == 0) { ph_resize(&t, t.cap * 2 + 1); s = ph_first_non_full(&t, ph_mix((uint64_t)keys[i])); }
        t.size++; t.growth--;
        t.full[s] = 1; t.key[s] = keys[i];
    }
    int m = 0;
    for (int i = 0; i < t.cap; i++) if (t.full[i]) out[m++] = t.key[i];
    free(t.full); free(t.key);
}

/* KssdIncrementalCluster (greedy.cpp:1736-1880, kssd = 1) and MinHashIncrementalCluster (:1963-2130, kssd = 0) with one
 * thread: sketches [0, start) are clustered, seed[0..n_seed) are their representatives in creation order; sketch
 * q >= start joins the representative of smallest distance -- the first such one in the iteration order of the candidate
 * table -- or becomes the next representative.  rep_of[q] is written for q >= start only.  Returns #representatives. */
typedef struct { uint32_t first_hit; int rep_idx; int common; } inc_touch;
static int cmp_inc_touch(const void *a, const void *b)
{
    const inc_touch *x = (const inc_touch *)a, *y = (const inc_touch *)b;
    if (x->first_hit != y->first_hit) return x->first_hit < y->first_hit ? -1 : 1;
    return x->rep_idx < y->rep_idx ? -1 : x->rep_idx > y->rep_idx;
}
static int greedy_incremental(const uint64_t *hashes, const uint64_t *off, int N, int k, int kssd, int containment,
                              double threshold, int start, const int32_t *seed, int n_seed, int32_t *rep_of)
{
    const size_t cap = (size_t)(N > 0 ? N : 1);
    int *reps = (int *)malloc(sizeof(int) * cap), nrep = 0;
    inc_touch *tc = (inc_touch *)malloc(sizeof(inc_touch) * cap);
    int32_t *ins = (int32_t *)malloc(4 * cap), *ord = (int32_t *)malloc(4 * cap);
    int *common_of = (int *)malloc(sizeof(int) * cap);
    for (int r = 0; r < n_seed; r++) reps[nrep++] = seed[r];
    const double radio = 2.0 * exp(threshold * k) - 1.0;              /* calculateMaxSizeRatio, greedy.cpp:162-173 */
    const double x = exp(-threshold * k);                             /* :1751-1752, :2000-2001 */
    const double jaccard_min = x / (2.0 - x);
    for (int q = start; q < N; q++) {
        const uint64_t *hq = hashes + off[q];
        const int sizeQry = (int)(off[q + 1] - off[q]);
        /* the index probe (:1774-1795, :2017-2024): query hashes in array order, postings in representative order */
        int nt = 0;
        for (int r = 0; r < nrep; r++) {
            const int id = reps[r];
            uint32_t fh;
            const int c = (int)oc_intersect_u64(hq, sizeQry, hashes + off[id], off[id + 1] - off[id], &fh);
            if (c > 0) { tc[nt].first_hit = fh; tc[nt].rep_idx = r; tc[nt].common = c; nt++; }
        }
        qsort(tc, nt, sizeof(inc_touch), cmp_inc_touch);              /* first-insertion order into candidate_counts */
        for (int t = 0; t < nt; t++) { ins[t] = tc[t].rep_idx; common_of[tc[t].rep_idx] = tc[t].common; }
        oc_phmap_iteration_order(ins, nt, ord);                       /* `for (const auto& pair : candidate_counts)` */
        double best_dist = 1.7976931348623157e308;
        int best = -1;
        for (int t = 0; t < nt; t++) {
            const int r = ord[t], common = common_of[r];
            const int sizeRef = (int)(off[reps[r] + 1] - off[reps[r]]);
            if (sizeRef == 0) continue;
            if (kssd) {                                               /* :1818-1821 */
                const double ratio = (double)sizeQry / sizeRef;
                if (ratio > radio || ratio < 1.0 / radio) continue;
            }
            int min_common_needed;                                    /* :1825, :2057-2063 */
            if (containment) {
                const int minSz = sizeQry < sizeRef ? sizeQry : sizeRef;
                min_common_needed = (int)(jaccard_min * minSz);
            } else {
                min_common_needed = (int)(jaccard_min * (sizeQry + sizeRef) / (1.0 + jaccard_min));
            }
            if (common < min_common_needed) continue;
            double jaccard;                                           /* calculate_mash_distance :72-119 / inline :2069-2085 */
            if (containment) {
                const int minSz = sizeQry < sizeRef ? sizeQry : sizeRef;
                jaccard = (double)common / minSz;
            } else {
                const int denom = sizeQry + sizeRef - common;
                if (denom == 0) continue;
                jaccard = (double)common / denom;
            }
            double dist;
            if (jaccard >= 1.0) dist = 0.0;
            else if (jaccard <= 0.0) dist = 1.0;
            else {
                dist = -log(2.0 * jaccard / (1.0 + jaccard)) / k;
                if (dist > 1.0) dist = 1.0;
            }
            if (dist <= threshold && dist < best_dist) { best_dist = dist; best = r; }   /* :1840-1848, :2087-2094 */
        }
        if (best != -1) rep_of[q] = reps[best];
        else { reps[nrep++] = q; rep_of[q] = q; }
    }
    free(reps); free(tc); free(ins); free(ord); free(common_of);
    return nrep;
}

int oc_greedy_incremental_kssd(const uint64_t *hashes, const uint64_t *off, int N, int k, double threshold,
                               int start, const int32_t *seed, int n_seed, int32_t *rep_of)
{
    return greedy_incremental(hashes, off, N, k, 1, 0, threshold, start, seed, n_seed, rep_of);
}

int oc_greedy_incremental_minhash(const uint64_t *hashes, const uint64_t *off, int N, int k, int containment,
                                  double threshold, int start, const int32_t *seed, int n_seed, int32_t *rep_of)
{
    return greedy_incremental(hashes, off, N, k, 0, containment, threshold, start, seed, n_seed, rep_of);
}

/* ------------------------------------------------- semantic B brute-force callers */
/* MinHash::jaccard as integers: common among the first sketch_size union elements and the
 * completed denominator (MinHash.cpp:982-1101, scalar form :1199-1229). */
uint32_t oc_common_B(const uint64_t *a, size_t na, const uint64_t *b, size_t nb,
                     uint64_t sketch_size, uint32_t *denom_out)
{
    uint64_t i = 0, j = 0, common = 0, denom = 0;
    while (denom < sketch_size && i < na && j < nb) {
        if (a[i] < b[j]) i++;
        else if (b[j] < a[i]) j++;
        else { i++; j++; common++; }
        denom++;
    }
    if (denom < sketch_size) {
        if (i < na) denom += na - i;
        if (j < nb) denom += nb - j;
        if (denom > sketch_size) denom = sketch_size;
    }
    if (denom_out) *denom_out = (uint32_t)denom;
    return (uint32_t)common;
}

/* MinHash::containDistance (MinHash.cpp:1102-1123) over containJaccard (:908-979) */
double oc_contain_distance(const uint64_t *a, size_t na, const uint64_t *b, size_t nb, int k)
{
    uint32_t common = oc_intersect_u64(a, na, b, nb, NULL);
    double j = (double)common / (double)(na < nb ? na : nb);
    if (j == 1.0) return 0;
    double d = -log(2 * j / (1. + j)) / k;
    if (d > 1) d = 1;
    return d;
}

static double pair_dist_B(const uint64_t *hashes, const uint64_t *off, int i, int j, int k,
                          uint64_t sketch_size, int containment)
{
    const uint64_t *a = hashes + off[i], *b = hashes + off[j];
    size_t na = off[i + 1] - off[i], nb = off[j + 1] - off[j];
    if (containment) return oc_contain_distance(a, na, b, nb, k);
    return oc_distance_B(oc_jaccard_B(a, na, b, nb, sketch_size), k);
}

/* modifyMST (MST.cpp:809-1018): every pair i<j gets MinHash::distance / containDistance, then
 * sort + Kruskal over the complete graph.  Returns N-1 edges. */
size_t oc_mst_bruteforce(const uint64_t *hashes, const uint64_t *off, int N, int k,
                         uint64_t sketch_size, int containment, oc_edge *out)
{
    size_t n = 0, cap = (size_t)N * (N > 0 ? N - 1 : 0) / 2 + 1;
    oc_edge *g = (oc_edge *)malloc(cap * sizeof(oc_edge));
    for (int i = 0; i < N; i++)
        for (int j = i + 1; j < N; j++) {
            g[n].pre = i; g[n].suf = j;
            g[n].dist = pair_dist_B(hashes, off, i, j, k, sketch_size, containment);
            n++;
        }
    qsort(g, n, sizeof(oc_edge), cmp_edge);
    int *p = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    for (int i = 0; i < N; i++) p[i] = i;
    size_t m = 0;
    for (size_t e = 0; e < n; e++) {
        int a = uf_find(p, g[e].pre), b = uf_find(p, g[e].suf);
        if (a != b) { p[a] = b; out[m++] = g[e]; }
    }
    free(p); free(g);
    return m;
}

/* greedyCluster (greedy.cpp:285-351) with one thread: genome j joins the nearest
 * representative with dist <= threshold (std::map<double,int>::insert keeps the FIRST rep
 * inserted for equal distances = earliest-created), else becomes a representative. */
int oc_greedy_bruteforce(const uint64_t *hashes, const uint64_t *off, int N, int k,
                         uint64_t sketch_size, int containment, double threshold, int32_t *rep_of)
{
    if (N == 0) return 0;
    int *reps = (int *)malloc(sizeof(int) * N), nrep = 0;
    reps[nrep++] = 0; rep_of[0] = 0;
    for (int j = 1; j < N; j++) {
        int best = -1;
        double best_d = 0;
        for (int r = 0; r < nrep; r++) {
            double d = pair_dist_B(hashes, off, reps[r], j, k, sketch_size, containment);
            if (d <= threshold && (best < 0 || d < best_d)) { best = reps[r]; best_d = d; }
        }
        if (best >= 0) rep_of[j] = best;
        else { reps[nrep++] = j; rep_of[j] = j; }
    }
    free(reps);
    return nrep;
}

/* ------------------------------------------------------------------ --dense epilogue */
/* the `if(!no_dense)` block of compute_minhash_mst (MST.cpp:1517-1530) over the candidate pairs
 * (common >= 1, size-ratio filter), then the cumulative pass (:1702-1712).  dense: span x N ints,
 * ani: 101 counters. */
void oc_dense_ani(const uint64_t *hashes, const uint64_t *off, int N, int k, int containment,
                  double threshold, int span, int32_t *dense, uint64_t *ani)
{
    int radio = oc_radio(threshold, k);
    double step = 1.0 / span;
    double *radius = (double *)malloc(sizeof(double) * span);
    for (int t = 0; t < span; t++) radius[t] = step * (double)t;
    memset(dense, 0, sizeof(int32_t) * (size_t)span * N);
    memset(ani, 0, sizeof(uint64_t) * 101);
    for (int i = 0; i < N; i++) {
        int size0 = (int)(off[i + 1] - off[i]);
        if (size0 == 0) continue;
        for (int j = 0; j < i; j++) {
            int size1 = (int)(off[j + 1] - off[j]);
            if (size1 == 0) continue;
            int common = (int)oc_intersect_u64(hashes + off[i], size0, hashes + off[j], size1, NULL);
            if (common == 0) continue;
            int mn = size0 < size1 ? size0 : size1, mx = size0 > size1 ? size0 : size1;
            if (mx > radio * mn) continue;
            double d = containment ? oc_aaf_dist_mst(common, size0, size1, k) : oc_mash_dist_A(common, size0, size1, k);
            int t0 = 0;                                   /* std::lower_bound(distRadius, ..., d) */
            while (t0 < span && radius[t0] < d) t0++;
            if (t0 < span) { dense[(size_t)t0 * N + i]++; dense[(size_t)t0 * N + j]++; }
            double tmpANI = 1.0 - d;
            int ANI = (int)(tmpANI * 100.0);
            if (ANI >= 101) ANI = 100;
            ani[ANI]++;
        }
    }
    for (int v = 0; v < N; v++) {
        int acc = 0;
        for (int t = 0; t < span; t++) { acc += dense[(size_t)t * N + v]; dense[(size_t)t * N + v] = acc; }
    }
    free(radius);
}

/* getNoiseNode (MST.cpp:189-211) per cluster of size > 1 with alpha = 2 at denseIndex = threshold / 0.01
 * (sub_command.cpp:3075-3092), then modifyForest (:87-106) and the BFS clusters: labels/order as
 * oc_clusters_from_mst.  forest = MST edges with dist <= threshold.  Returns #clusters. */
int oc_remove_noise_clusters(const oc_edge *mst, size_t n_edges, int N, double threshold, const int32_t *dense,
                             int span, int32_t *labels, int32_t *order)
{
    int32_t *lab0 = (int32_t *)malloc(sizeof(int32_t) * (N + 1)), *ord0 = (int32_t *)malloc(sizeof(int32_t) * (N + 1));
    int nc = oc_clusters_from_mst(mst, n_edges, N, threshold, lab0, ord0);
    int alpha = 2;
    int denseIndex = threshold / 0.01;
    char *noise = (char *)calloc(N + 1, 1);
    int *vals = (int *)malloc(sizeof(int) * (N + 1));
    int pos = 0;
    for (int c = 0; c < nc; c++) {
        int start = pos, m = 0;
        while (pos < N && lab0[ord0[pos]] == c) { vals[m++] = dense[(size_t)denseIndex * N + ord0[pos]]; pos++; }
        if (m == 1) continue;
        /* Q1 of the sorted density values */
        int *sorted = (int *)malloc(sizeof(int) * m);
        memcpy(sorted, vals, sizeof(int) * m);
        for (int a = 1; a < m; a++) { int x = sorted[a], b = a - 1; while (b >= 0 && sorted[b] > x) { sorted[b + 1] = sorted[b]; b--; } sorted[b + 1] = x; }
        int denseQ1 = sorted[m / 4];
        int thr = alpha;
        if (denseQ1 - 1 < thr) thr = denseQ1 - 1;
        if (thr < 0) thr = 0;
        for (int a = 0; a < m; a++) if (vals[a] <= thr) noise[ord0[start + a]] = 1;
        free(sorted);
    }
    oc_edge *kept = (oc_edge *)malloc(sizeof(oc_edge) * (n_edges + 1));
    size_t nk = 0;
    for (size_t e = 0; e < n_edges; e++)
        if (mst[e].dist <= threshold && !noise[mst[e].pre] && !noise[mst[e].suf]) kept[nk++] = mst[e];
    int res = oc_clusters_from_mst(kept, nk, N, threshold, labels, order);
    free(lab0); free(ord0); free(noise); free(vals); free(kept);
    (void)span;
    return res;
}
