/*
 * ref_harness.cpp -- thin driver around the UNMODIFIED reference library
 * (RabbitSketch, compiled from /root/reference by oracle/Makefile into oracle/_ref/).
 * TEST / BASELINE INFRASTRUCTURE ONLY: used to validate oracle/tclust_oracle.c, to
 * generate tests/golden/, and as the "reference" CPU baseline timed by bench.py.
 * It contains no reference source; it only calls the reference's public API
 * (Sketch::MinHash, RabbitSketch/src/Sketch.h:55-164).
 *
 *   ref_harness sketch <k> <s> <fasta[.gz]>...      one sketch per FILE (as sketchFiles,
 *                                                   src/SketchInfo.cpp:865-992); binary out:
 *                                                   per file  u64 n, u64 h[n]
 *   ref_harness pairs <hash.sketch> <threads> <nmax> [contain]
 *                                                   all-pairs MinHash::distance (semantic B,
 *                                                   MinHash.cpp:1125-1153) over the first nmax
 *                                                   sketches; prints JSON with pairs/s
 *   ref_harness update <k> <s> <threads> <len> <ngenomes>
 *                                                   time MinHash::update on synthetic genomes
 *   ref_harness dist <hash.sketch> <i> <j>          jaccard/distance of one pair (semantic B)
 *   ref_harness pairsdump <hash.sketch> <nmax>      binary: for i, for j<i: double jaccard(), distance(),
 *                                                   containDistance()  (MinHash.cpp:982-1153, 908-979)
 *   ref_harness mstb <hash.sketch> <nmax> <threads> [contain]
 *                                                   the reference's brute-force MST modifyMST
 *                                                   (src/MST.cpp:809-1018); binary: u64 n, {i32,i32,f64}[n]
 *   ref_harness msti <folder> <threads> <threshold> <out.bin>
 *                                                   the reference's default MST path on a --presketched folder:
 *                                                   loadSketches + MinHashInvertedIndex (loaded from the folder's cache or
 *                                                   built) + compute_minhash_mst (src/MST.cpp:1290-1737, incl. its final
 *                                                   Kruskal); out.bin: u64 n, {i32,i32,f64}[n].  clust-mst --presketched
 *                                                   itself never writes edge.mst, so this is how the full-size MST weights
 *                                                   of the unmodified reference are obtained (bench.py parity)
 *   ref_harness phorder <keys.bin> <out.bin>        the iteration order of the reference's candidate table: the int32 keys of
 *                                                   keys.bin go, in order, through `phmap::flat_hash_map<int,int> m;
 *                                                   m.reserve(1000); m[key]++` (KssdIncrementalCluster / MinHashIncrementalCluster,
 *                                                   src/greedy.cpp:1771-1803, 2014-2036, with the reference's own phmap.h);
 *                                                   out.bin: the keys in iteration order
 *   ref_harness kssdinc <kssd folder> <n_old> <threshold>
 *   ref_harness minhashinc <folder> <n_old> <threshold>
 *                                                   the reference's incremental greedy rules on a pre-sketched folder: the
 *                                                   first n_old sketches go through KssdInitialClusterWithState /
 *                                                   MinHashInitialClusterWithState, the rest through KssdIncrementalCluster /
 *                                                   MinHashIncrementalCluster (src/greedy.cpp:902-957, 1736-1880, 1904-2130), one
 *                                                   thread.  text: line 1 = folder index of every sketch in state order, line 2 =
 *                                                   representative ids, then one cluster per line ("-" = empty)
 *   ref_harness greedyb <hash.sketch> <nmax> <threads> <threshold> [contain]
 *                                                   the reference's brute-force greedyCluster
 *                                                   (src/greedy.cpp:285-351); text: one cluster per line
 */
#include "Sketch.h"
#include "MST.h"
#include "greedy.h"
#include "Sketch_IO.h"
#include "SketchInfo.h"
#include <zlib.h>
#include "kseq.h"
#include <omp.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

KSEQ_INIT(gzFile, gzread)

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Folder { int func_id, k; bool contain; int param; std::vector<std::vector<uint64_t>> sk; };

static bool load_hash_sketch(const char *path, size_t nmax, Folder &f)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return false;
    if (fread(&f.func_id, 4, 1, fp) != 1 || fread(&f.k, 4, 1, fp) != 1 ||
        fread(&f.contain, 1, 1, fp) != 1 || fread(&f.param, 4, 1, fp) != 1) { fclose(fp); return false; }
    size_t n;
    while (f.sk.size() < nmax && fread(&n, sizeof(size_t), 1, fp) == 1) {
        std::vector<uint64_t> h(n);
        if (n && fread(h.data(), 8, n, fp) != n) break;
        f.sk.push_back(std::move(h));
    }
    fclose(fp);
    return true;
}

static int cmd_sketch(int argc, char **argv)
{
    int k = atoi(argv[2]), s = atoi(argv[3]);
    for (int a = 4; a < argc; a++) {
        gzFile fp = gzopen(argv[a], "r");
        if (!fp) { fprintf(stderr, "cannot open %s\n", argv[a]); return 1; }
        kseq_t *ks = kseq_init(fp);
        Sketch::MinHash mh(k, s);
        while (kseq_read(ks) >= 0) mh.update(ks->seq.s);
        std::vector<uint64_t> h = mh.storeMinHashes();
        uint64_t n = h.size();
        fwrite(&n, 8, 1, stdout);
        fwrite(h.data(), 8, n, stdout);
        kseq_destroy(ks);
        gzclose(fp);
    }
    return 0;
}

static int cmd_pairs(int argc, char **argv)
{
    Folder f;
    int threads = atoi(argv[3]);
    size_t nmax = strtoull(argv[4], 0, 10);
    bool contain = argc > 5 && !strcmp(argv[5], "contain");
    if (!load_hash_sketch(argv[2], nmax, f)) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    size_t N = f.sk.size();
    std::vector<Sketch::MinHash *> mh(N);
    for (size_t i = 0; i < N; i++) {
        mh[i] = new Sketch::MinHash(f.k, f.param);
        mh[i]->loadMinHashes(f.sk[i]);
        mh[i]->getSketchSize();
    }
    double sum = 0;
    double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8) reduction(+ : sum)
    for (long i = 0; i < (long)N; i++)
        for (size_t j = 0; j < (size_t)i; j++)
            sum += contain ? mh[i]->containDistance(mh[j]) : mh[i]->distance(mh[j]);
    double t1 = now_s();
    double pairs = (double)N * (N - 1) / 2;
    printf("{\"n\": %zu, \"pairs\": %.0f, \"seconds\": %.6f, \"pairs_per_s\": %.1f, \"threads\": %d, \"dist_sum\": %.10f}\n",
           N, pairs, t1 - t0, pairs / (t1 - t0), threads, sum);
    return 0;
}

static int cmd_update(int argc, char **argv)
{
    int k = atoi(argv[2]), s = atoi(argv[3]), threads = atoi(argv[4]);
    size_t len = strtoull(argv[5], 0, 10);
    int ng = atoi(argv[6]);
    std::vector<std::string> g(ng);
    for (int i = 0; i < ng; i++) {
        uint64_t x = 0x9E3779B97F4A7C15ULL * (i + 1);
        g[i].resize(len);
        for (size_t p = 0; p < len; p++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            g[i][p] = "ACGT"[x & 3];
        }
    }
    uint64_t chk = 0;
    double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic) reduction(^ : chk)
    for (int i = 0; i < ng; i++) {
        Sketch::MinHash mh(k, s);
        mh.update(&g[i][0]);
        std::vector<uint64_t> h = mh.storeMinHashes();
        for (uint64_t v : h) chk ^= v;
    }
    double t1 = now_s();
    printf("{\"genomes\": %d, \"len\": %zu, \"seconds\": %.6f, \"genomes_per_s\": %.3f, \"mbp_per_s\": %.3f, \"threads\": %d, \"chk\": %llu}\n",
           ng, len, t1 - t0, ng / (t1 - t0), ng * (double)len / 1e6 / (t1 - t0), threads, (unsigned long long)chk);
    return 0;
}

static int cmd_dist(int argc, char **argv)
{
    Folder f;
    if (!load_hash_sketch(argv[2], (size_t)-1, f)) return 1;
    int i = atoi(argv[3]), j = atoi(argv[4]);
    Sketch::MinHash a(f.k, f.param), b(f.k, f.param);
    a.loadMinHashes(f.sk[i]); b.loadMinHashes(f.sk[j]);
    printf("%.17g %.17g %.17g\n", a.jaccard(&b), a.distance(&b), a.containDistance(&b));
    return 0;
}

static int cmd_pairsdump(int argc, char **argv)
{
    Folder f;
    if (!load_hash_sketch(argv[2], strtoull(argv[3], 0, 10), f)) return 1;
    size_t N = f.sk.size();
    std::vector<Sketch::MinHash *> mh(N);
    for (size_t i = 0; i < N; i++) { mh[i] = new Sketch::MinHash(f.k, f.param); mh[i]->loadMinHashes(f.sk[i]); }
    for (size_t i = 0; i < N; i++)
        for (size_t j = 0; j < i; j++) {
            double v[3] = {mh[i]->jaccard(mh[j]), mh[i]->distance(mh[j]), mh[i]->containDistance(mh[j])};
            fwrite(v, 8, 3, stdout);
        }
    return 0;
}

static void make_infos(const Folder &f, bool contain, std::vector<SketchInfo> &v)
{
    v.resize(f.sk.size());
    for (size_t i = 0; i < f.sk.size(); i++) {
        v[i].id = (int)i;
        v[i].isContainment = contain;
        v[i].totalSeqLength = 0;
        v[i].minHash = new Sketch::MinHash(f.k, f.param);
        v[i].minHash->loadMinHashes(const_cast<std::vector<uint64_t> &>(f.sk[i]));
    }
}

static int cmd_mstb(int argc, char **argv)
{
    Folder f;
    if (!load_hash_sketch(argv[2], strtoull(argv[3], 0, 10), f)) return 1;
    int threads = atoi(argv[4]);
    bool contain = argc > 5 && !strcmp(argv[5], "contain");
    std::vector<SketchInfo> v;
    make_infos(f, contain, v);
    int **dense = nullptr; uint64_t *ani = nullptr;
    std::vector<EdgeInfo> mst = modifyMST(v, 0, 0, threads, true, dense, 0, ani);
    uint64_t n = mst.size();
    fwrite(&n, 8, 1, stdout);
    for (const EdgeInfo &e : mst) { fwrite(&e.preNode, 4, 1, stdout); fwrite(&e.sufNode, 4, 1, stdout); fwrite(&e.dist, 8, 1, stdout); }
    return 0;
}

static int cmd_msti(int argc, char **argv)
{
    const std::string folder = argv[2];
    const int threads = atoi(argv[3]);
    const double thr = atof(argv[4]);
    std::vector<SketchInfo> v;
    int func_id = 0;
    const double t0 = now_s();
    loadSketches(folder, threads, v, func_id);
    if (func_id != 0 || v.empty() || v[0].minHash == nullptr) { fprintf(stderr, "msti: not a MinHash folder\n"); return 1; }
    int fid, k = 21, cc, ss, hk, hs, dr;
    bool contain = false;
    read_sketch_parameters(folder, fid, k, contain, cc, ss, hk, hs, dr);
    MinHashInvertedIndex idx;
    if (!idx.load_from_file(folder)) {
        for (size_t i = 0; i < v.size(); i++) {
            v[i].id = (int)i;
            std::vector<uint64_t> h = v[i].minHash->storeMinHashes();
            for (uint64_t x : h) idx.hash_map[x].push_back((uint32_t)i);
        }
    }
    const double t1 = now_s();
    int **dense = nullptr; uint64_t *ani = nullptr;
    std::vector<EdgeInfo> mst = compute_minhash_mst(v, 0, true, contain, threads, dense, 0, ani, thr, k, &idx);
    const double t2 = now_s();
    FILE *fo = fopen(argv[5], "wb");
    if (!fo) { fprintf(stderr, "msti: cannot write %s\n", argv[5]); return 1; }
    uint64_t n = mst.size();
    fwrite(&n, 8, 1, fo);
    for (const EdgeInfo &e : mst) { fwrite(&e.preNode, 4, 1, fo); fwrite(&e.sufNode, 4, 1, fo); fwrite(&e.dist, 8, 1, fo); }
    fclose(fo);
    printf("{\"edges\": %llu, \"load_s\": %.3f, \"mst_s\": %.3f, \"threads\": %d}\n", (unsigned long long)n, t1 - t0, t2 - t1, threads);
    return 0;
}

static int cmd_greedyb(int argc, char **argv)
{
    Folder f;
    if (!load_hash_sketch(argv[2], strtoull(argv[3], 0, 10), f)) return 1;
    int threads = atoi(argv[4]);
    double thr = atof(argv[5]);
    bool contain = argc > 6 && !strcmp(argv[6], "contain");
    std::vector<SketchInfo> v;
    make_infos(f, contain, v);
    std::vector<std::vector<int>> cl = greedyCluster(v, 0, thr, threads);
    for (const auto &c : cl) {
        for (size_t i = 0; i < c.size(); i++) printf(i ? " %d" : "%d", c[i]);
        printf("\n");
    }
    return 0;
}

template <typename State>
static void print_state(const State &st, const std::vector<std::vector<int>> &cl)
{
    for (size_t i = 0; i < st.all_sketches.size(); i++) printf(i ? " %d" : "%d", st.all_sketches[i].id);
    printf("\n");
    for (size_t i = 0; i < st.representative_ids.size(); i++) printf(i ? " %d" : "%d", st.representative_ids[i]);
    printf("\n");
    for (const auto &c : cl) {
        if (c.empty()) printf("-");
        for (size_t i = 0; i < c.size(); i++) printf(i ? " %d" : "%d", c[i]);
        printf("\n");
    }
}

static int cmd_kssdinc(int argc, char **argv)
{
    std::vector<KssdSketchInfo> all;
    KssdParameters info;
    loadKssdSketches(argv[2], 1, all, info);
    const size_t n_old = strtoull(argv[3], 0, 10);
    const double thr = atof(argv[4]);
    if (n_old == 0 || n_old > all.size()) { fprintf(stderr, "kssdinc: bad n_old\n"); return 1; }
    for (size_t i = 0; i < all.size(); i++) all[i].id = (int)i;
    std::vector<KssdSketchInfo> pre(all.begin(), all.begin() + n_old), add(all.begin() + n_old, all.end());
    const int kmer_size = info.half_k * 2;
    KssdClusterState st = KssdInitialClusterWithState(pre, info, thr, 1, kmer_size);
    std::vector<std::vector<int>> cl = KssdIncrementalCluster(st, add, 1);
    print_state(st, cl);
    return 0;
}

static int cmd_minhashinc(int argc, char **argv)
{
    std::vector<SketchInfo> all;
    int func_id = 0;
    loadSketches(argv[2], 1, all, func_id);
    const size_t n_old = strtoull(argv[3], 0, 10);
    const double thr = atof(argv[4]);
    if (n_old == 0 || n_old > all.size() || func_id != 0) { fprintf(stderr, "minhashinc: bad input\n"); return 1; }
    for (size_t i = 0; i < all.size(); i++) all[i].id = (int)i;
    std::vector<SketchInfo> pre(all.begin(), all.begin() + n_old), add(all.begin() + n_old, all.end());
    const int kmer_size = pre[0].minHash->getKmerSize();
    const int sketch_size = pre[0].minHash->getSketchSize();
    const bool contain = pre[0].isContainment;
    MinHashClusterState st = MinHashInitialClusterWithState(pre, thr, 1, kmer_size, sketch_size, contain);
    std::vector<std::vector<int>> cl = MinHashIncrementalCluster(st, add, 1);
    print_state(st, cl);
    return 0;
}

static int cmd_phorder(int argc, char **argv)
{
    FILE *fi = fopen(argv[2], "rb");
    if (!fi) { perror(argv[2]); return 1; }
    std::vector<int> keys;
    int k;
    while (fread(&k, 4, 1, fi) == 1) keys.push_back(k);
    fclose(fi);
    phmap::flat_hash_map<int, int> m;
    m.reserve(1000);
    for (int x : keys) m[x]++;
    FILE *fo = fopen(argv[3], "wb");
    if (!fo) { perror(argv[3]); return 1; }
    for (const auto &pr : m) fwrite(&pr.first, 4, 1, fo);
    fclose(fo);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 4 && !strcmp(argv[1], "phorder")) return cmd_phorder(argc, argv);
    if (argc >= 5 && !strcmp(argv[1], "kssdinc")) return cmd_kssdinc(argc, argv);
    if (argc >= 5 && !strcmp(argv[1], "minhashinc")) return cmd_minhashinc(argc, argv);
    if (argc >= 4 && !strcmp(argv[1], "pairsdump")) return cmd_pairsdump(argc, argv);
    if (argc >= 5 && !strcmp(argv[1], "mstb")) return cmd_mstb(argc, argv);
    if (argc >= 6 && !strcmp(argv[1], "msti")) return cmd_msti(argc, argv);
    if (argc >= 6 && !strcmp(argv[1], "greedyb")) return cmd_greedyb(argc, argv);
    if (argc >= 5 && !strcmp(argv[1], "sketch")) return cmd_sketch(argc, argv);
    if (argc >= 5 && !strcmp(argv[1], "pairs")) return cmd_pairs(argc, argv);
    if (argc >= 7 && !strcmp(argv[1], "update")) return cmd_update(argc, argv);
    if (argc >= 5 && !strcmp(argv[1], "dist")) return cmd_dist(argc, argv);
    fprintf(stderr, "usage: see header of ref_harness.cpp\n");
    return 2;
}
