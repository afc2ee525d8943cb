// tclust_cli.cpp -- host side of the drop-in: `clust-mst` and `clust-greedy` (-DGREEDY_CLUST) with the
// reference's in-scope flags, parameter tuning, on-disk formats and cluster-file output, calling the CUDA
// hot paths through the C ABI of include/tclust.h.  There is no CPU implementation of the hot paths here.
//
// Reference behaviour mirrored (paths relative to the reference root):
//   options / dispatch        src/main.cpp:75-129,200,255-289,639-671
//   calSize, tune_parameters  src/SketchInfo.cpp:438-552, src/sub_command.cpp:2383-2467
//   sketchFiles (list mode)   src/SketchInfo.cpp:865-992      (-c size rule :892-921, -m filter :963)
//   sketchSequences (-i file) src/SketchInfo.cpp:205-272,741-763; RabbitFX/io/Formater.cpp:87-129
//   saveSketches/loadSketches src/Sketch_IO.cpp:88-226,284-353,477-560
//   saveMST/loadMST           src/MST_IO.cpp:47-70,200-217
//   forest / BFS clusters     src/MST.cpp:77-85,109-142
//   printResult               src/MST_IO.cpp:136-179
//   greedy cluster order      src/greedy.cpp:1321-1336,1358-1367
//   --fast (KSSD, clust-mst)  src/main.cpp:603-636, src/sub_command.cpp:1934-2010,2317-2382 (tuning),
//                             src/SketchInfo.cpp:994-1252 (list-mode driver), 60-102 (shuffled_dim),
//                             src/Sketch_IO.cpp:36-86,136-167,228-282 (kssd.*.sketch), src/SketchInfo.cpp:1379-1467
//                             (kssd.sketch.index/.dict), src/MST_IO.cpp:181-198 (kssd.info.mst + edge.mst)
// FASTA feeder: files are decompressed/parsed on host threads into page-locked batch buffers
// (tc_host_alloc); a batch is sketched by a background thread (tc_sketch_* pipelines its H2D copies) while the
// next batch is being parsed.
//   --dense                   src/MST.cpp:1517-1530,1702-1712 (tc_dense_ani), src/sub_command.cpp:3070-3100 (noise pass)
//   tree / linkage output     src/MST.cpp:1088-1147,1235-1284, src/MST_IO.cpp:263-375
//   --append (classic)        src/sub_command.cpp:22-96 (greedy: brute-force greedyCluster), :1578-1757 (mst: new-vs-all
//                             pass + Kruskal over old MST + new forest), :1317-1477 (clust-mst --fast: the same over
//                             KSSD sketches, append_clust_mst_fast without a saved state)
// Out of scope: --db/--save-rep state files, --auto-threshold/--stability analytics, dbscan/leiden,
// --fast with a single multi-FASTA input (the reference's -i KSSD consumer is the
// quirky variant).
#include "../../include/tclust.h"
#include <zlib.h>
#include <sys/stat.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <future>
#include <iostream>
#include <queue>
#include <string>
#include <mutex>
#include <memory>
#include <functional>
#include <condition_variable>
#include <deque>
#include <thread>
#include <vector>
#include <atomic>

using std::cerr;
using std::endl;
using std::string;
using std::vector;

bool path_is_gz(const string &p) { return p.size() >= 2 && p.compare(p.size() - 2, 2, "gz") == 0; }

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Genome {          // SketchInfo (src/SketchInfo.h:23-36) without the owning sketch pointers
    string file_name;    // list mode
    string name, comment;  // first record (list mode) or the record (single-file mode)
    int strand = 0;
    uint64_t total_len = 0;   // list mode: sum of record lengths; single-file mode: record length
    vector<uint64_t> hashes;
    int id = 0;
    uint32_t ssize = 0;       // the sketch-size argument its MinHash object was built with (getSketchSize(): state files)
    bool from_state = false;  // rebuilt from a KSSD cluster_state.bin: no record names, no per-record length (KssdClusterState::load)
};

struct Options {
    string input = "genome.fna", output = "result.out", folder;
    int threads = (int)std::thread::hardware_concurrency();
    bool by_file = false, containment = false, jaccard = false, set_kmer = false, no_save = false, fast = false;
    bool has_output = false, has_input = false, presketched = false, premsted = false, has_threshold = false;
    bool append = false, save_rep = false;
    bool dense = false, newick = false, phylip = false, nexus = false, linkage = false;
    double threshold = 0.05;
    int kmer = 19, sketch_size = 1000, compress = 1000, drlevel = 3, device = 0;
    bool threshold_only = false;   // --threshold-only: clusters at -d without building / saving the MST (tc_cluster_threshold)
    int gpus = 1;                 // --gpus G: devices device .. device+G-1 build the MST together (multi-GPU group of the library)
    uint64_t min_len = 10000;
};

[[noreturn]] void die(const string &msg) { cerr << "ERROR: " << msg << endl; exit(1); }
void tc_or_die(int rc, const char *where) { if (rc) die(string(where) + ", " + tc_last_error()); }

// ------------------------------------------------------------------------------------ FASTA
struct Record { string name, comment; bool has_comment = false; uint64_t begin = 0, end = 0; };

bool read_whole_gz(const string &path, string &out)
{
    {   // not gzip (magic 1f 8b): one read of the whole file instead of zlib's transparent copy loop
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) return false;
        unsigned char magic[2] = {0, 0};
        const size_t got = fread(magic, 1, 2, f);
        if (!(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
            struct stat sb;
            out.clear();
            if (fstat(fileno(f), &sb) == 0 && S_ISREG(sb.st_mode)) {
                out.resize((size_t)sb.st_size);
                rewind(f);
                const size_t n = sb.st_size ? fread(&out[0], 1, (size_t)sb.st_size, f) : 0;
                out.resize(n);
                fclose(f);
                return true;
            }
        }
        fclose(f);
    }
    gzFile fp = gzopen(path.c_str(), "r");
    if (!fp) return false;
    gzbuffer(fp, 1 << 20);
    out.clear();
    char buf[1 << 16];
    int n;
    while ((n = gzread(fp, buf, sizeof buf)) > 0) out.append(buf, n);
    gzclose(fp);
    return true;
}

// Parse FASTA text; sequence bytes (line breaks removed) are appended to `seq`, records to `recs`.
// kseq_mode: name ends at the first whitespace, comment = rest of the header line (kseq.h);
// otherwise RabbitFX's rule: name ends at the first ' ' (Formater.cpp:104-106).
void parse_fasta(const string &text, bool kseq_mode, vector<uint8_t> &seq, vector<Record> &recs)
{
    size_t p = 0, n = text.size();
    seq.reserve(seq.size() + n);                            // one allocation: the sequence is the text minus headers and line ends
    while (p < n && text[p] != '>') {                      // skip anything before the first header
        const char *nl = (const char *)memchr(text.data() + p, '\n', n - p);
        if (!nl) return;
        p = nl - text.data() + 1;
    }
    while (p < n) {
        size_t eol = p;
        { const char *nl = (const char *)memchr(text.data() + p, '\n', n - p); eol = nl ? (size_t)(nl - text.data()) : n; }
        size_t hend = eol;
        if (hend > p && text[hend - 1] == '\r') hend--;
        Record r;
        size_t q = p + 1;
        if (kseq_mode) {
            while (q < hend && !isspace((unsigned char)text[q])) q++;
            r.name.assign(text, p + 1, q - p - 1);
            if (q < hend) { r.comment.assign(text, q + 1, hend - q - 1); r.has_comment = true; }
        } else {
            while (q < hend && text[q] != ' ') q++;
            r.name.assign(text, p + 1, q - p - 1);
            if (q < hend) { r.comment.assign(text, q + 1, hend - q - 1); r.has_comment = true; }
        }
        r.begin = seq.size();
        p = eol < n ? eol + 1 : n;
        while (p < n && text[p] != '>') {
            const char *nl = (const char *)memchr(text.data() + p, '\n', n - p);
            size_t le = nl ? (size_t)(nl - text.data()) : n;
            size_t e = le;
            if (e > p && text[e - 1] == '\r') e--;
            seq.insert(seq.end(), text.begin() + p, text.begin() + e);
            p = le < n ? le + 1 : n;
        }
        r.end = seq.size();
        recs.push_back(std::move(r));
    }
}

// size of the genome file as the reference sees it (SketchInfo.cpp:453-470,892-915): gz trailer or file bytes
bool file_size_like_reference(const string &path, uint64_t &size_u64, int &size_int)
{
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) return false;
    if (path.size() >= 2 && path.compare(path.size() - 2, 2, "gz") == 0) {
        int v = 0;
        fseek(fp, -4, SEEK_END);
        if (fread(&v, sizeof(int), 1, fp) != 1) v = 0;
        size_int = v; size_u64 = (uint64_t)(int64_t)v;
    } else {
        fseek(fp, 0, SEEK_END);
        long v = ftell(fp);
        size_int = (int)v; size_u64 = (uint64_t)v;
    }
    fclose(fp);
    return true;
}

vector<string> read_list(const string &path)
{
    std::ifstream ifs(path);
    if (!ifs) die("cannot open the inputFile: " + path);
    vector<string> v;
    string line;
    while (getline(ifs, line)) v.push_back(line);
    return v;
}

uint64_t g_input_bytes = 0;              // total size of the input genomes (cal_size), sizes the feeder's buffers

// ------------------------------------------------------------------------ parameter tuning
void cal_size(const Options &o, uint64_t &max_size, uint64_t &min_size, uint64_t &avg_size)
{
    max_size = 0; min_size = (uint64_t)(int64_t)(int)(1u << 31); avg_size = 0;   // reference: `1 << 31` as int
    uint64_t total = 0; int number = 0, bad = 0;
    if (o.by_file) {
        for (const string &line : read_list(o.input)) {
            uint64_t cur = 0; int dummy = 0;
            if (path_is_gz(line)) file_size_like_reference(line, cur, dummy);
            else { struct stat sb; cur = stat(line.c_str(), &sb) == 0 ? (uint64_t)sb.st_size : 0; }
            if (cur < o.min_len) { bad++; continue; }
            max_size = std::max(max_size, cur); min_size = std::min(min_size, cur); total += cur; number++;
        }
    } else {
        string text;
        if (!read_whole_gz(o.input, text)) die("cannot open the genome file, " + o.input);
        vector<uint8_t> seq; vector<Record> recs;
        parse_fasta(text, false, seq, recs);
        max_size = 1;                                        // reference consumer init (SketchInfo.cpp:487)
        for (const Record &r : recs) {
            uint64_t len = r.end - r.begin;
            if (len < o.min_len) { bad++; continue; }
            max_size = std::max(max_size, len); min_size = std::min(min_size, len); total += len; number++;
        }
    }
    if (number == 0) die("calSize(), no genome passes the minimum length filter");
    avg_size = total / number;
    g_input_bytes = total;
    int total_number = number + bad;
    cerr << "\t===the genome number for clustering is: " << number << endl;
    cerr << "\t===the genome number below the minimum genome length threshold is: " << bad << endl;
    cerr << "\t===the total genome number is: " << total_number << endl;
    if ((double)bad / total_number >= 0.2)
        fprintf(stderr, "Warning: there are %d poor quality (length < %ld) genome assemblies in the total %d genome assemblied.\n",
                bad, (long)o.min_len, total_number);
    cerr << "\t===the totalSize is: " << total << endl;
    cerr << "\t===the maxSize is: " << max_size << endl;
    cerr << "\t===the minSize is: " << min_size << endl;
    cerr << "\t===the averageSize is: " << avg_size << endl;
}

// the two steps tune_parameters (src/sub_command.cpp:2383-2467) and tune_kssd_parameters (:2317-2382) share: the k-mer size
// rule (k is replaced when it is too small for the largest genome or more than 3 above the recommendation) and
// the check of the threshold against the largest distance the sketches can resolve.  The stderr lines are the reference's.
void tune_kmer_size(Options &o, uint64_t max_size)
{
    const double warning_rate = 0.01, recommend_rate = 0.0001;
    const int recommended = ceil(log(max_size * (1 - recommend_rate) / recommend_rate) / log(4));
    const int warning = ceil(log(max_size * (1 - warning_rate) / warning_rate) / log(4));
    if (!o.set_kmer) o.kmer = recommended;
    else if (o.kmer < warning) {
        cerr << "the kmerSize " << o.kmer << " is too small for the maximum genome size of " << max_size << endl;
        cerr << "replace the kmerSize to the: " << recommended << " for reducing the random collision of kmers" << endl;
        o.kmer = recommended;
    } else if (o.kmer > recommended + 3) {
        cerr << "the kmerSize " << o.kmer << " maybe too large for the maximum genome size of " << max_size << endl;
        cerr << "replace the kmerSize to the " << recommended << " for increasing the sensitivity of genome comparison" << endl;
        o.kmer = recommended;
    }
}

bool check_threshold_and_report(const Options &o, double min_jaccard)
{
    const double max_dist = min_jaccard >= 1.0 ? 1.0 : -1.0 / o.kmer * log(2 * min_jaccard / (1.0 + min_jaccard));
    cerr << "-----the max recommand distance threshold is: " << max_dist << endl;
    if (o.threshold > max_dist) {
        cerr << "ERROR: tune_parameters(), the threshold: " << o.threshold << " is out of the valid distance range estimated by Mash distance or AAF distance" << endl;
        cerr << "Please set a distance threshold with -d option" << endl;
        return false;
    }
    cerr << (o.by_file ? "-----sketch by file!" : "-----sketch by sequence!") << endl;
    cerr << "-----the kmerSize is: " << o.kmer << endl;
    cerr << "-----the thread number is: " << o.threads << endl;
    cerr << "-----the threshold is: " << o.threshold << endl;
    return true;
}

bool tune_parameters(Options &o)
{
    uint64_t max_size, min_size, avg_size;
    cal_size(o, max_size, min_size, avg_size);
    if (o.containment && o.jaccard) {
        cerr << "ERROR: tune_parameters(), conflict distance measurement of Mash distance (fixed-sketch-size) and AAF distance (variable-sketch-size) " << endl;
        return false;
    }
#ifdef GREEDY_CLUST
    if (!o.containment && !o.jaccard) { o.compress = avg_size / 1000; o.containment = true; }
    else if (o.containment && avg_size / o.compress < 10) {
        cerr << "the containCompress " << o.compress << " is too large and the sketch size is too small" << endl;
        o.compress = avg_size / 1000;
        cerr << "set the containCompress to: " << o.compress << endl;
    }
#endif
    tune_kmer_size(o, max_size);
    if (!check_threshold_and_report(o, !o.containment ? 1.0 / o.sketch_size : 1.0 / (min_size / o.compress))) return false;
    if (o.containment) cerr << "-----use the AAF distance (variable-sketch-size), the sketchSize is in proportion with 1/" << o.compress << endl;
    else cerr << "-----use the Mash distance (fixed-sketch-size), the sketchSize is: " << o.sketch_size << endl;
    return true;
}

struct KssdParams { int id = 0, half_k = 0, half_subk = 0, drlevel = 0, genome_number = 0; };   // SketchInfo.h:52-58
KssdParams g_kssd;                       // set by sketch_files when o.fast
vector<int32_t> g_shuffled_dim;          // generate_shuffle_dim table (host only: glibc rand)

// tune_kssd_parameters (src/sub_command.cpp:2317-2382)
bool tune_kssd_parameters(Options &o)
{
    uint64_t max_size, min_size, avg_size;
    cal_size(o, max_size, min_size, avg_size);
    const int compression = 1 << (4 * o.drlevel);
    const int sketch_size = avg_size / compression;
    tune_kmer_size(o, max_size);
    if (!check_threshold_and_report(o, !o.containment ? 1.0 / sketch_size : 1.0 / (min_size / compression))) return false;
    if (o.containment) cerr << "-----use the AAF distance, the sketchSize is approximately in proportion with 1/" << compression << endl;
    else cerr << "-----use the Mash distance, the sketchSize is about: " << sketch_size << endl;
    return true;
}

// generate_shuffle_dim (src/SketchInfo.cpp:60-102) lives behind the ABI (host-only: glibc rand stream)
void make_shuffled_dim(int half_subk)
{
    g_shuffled_dim.resize((size_t)1 << (4 * half_subk));
    tc_or_die(tc_kssd_shuffled_dim(half_subk, g_shuffled_dim.data()), "shuffled_dim");
}

// ---------------------------------------------------------------------------- sketching
// growable byte buffer of a sketch batch: page-locked (tc_host_alloc: the H2D copies of tc_sketch_* then overlap the
// kernels and run at PCIe speed) when the buffers are reused for many batches, plain memory when the whole input is
// one or two batches -- page-locking costs ~0.4 ms per MB (470 ms for the 2 x 576 MB of BASELINE configs[0], ten times
// the copy it would speed up) and needs the CUDA context, which a small run would otherwise not wait for yet
struct PinnedBuf {
    uint8_t *p = nullptr; size_t cap = 0, len = 0; bool pinned = true;
    void reserve(size_t need)
    {
        if (need <= cap) return;
        // page-locking is slow: grow geometrically and let the feeder pre-size the buffer
        const size_t nc = std::max<size_t>(std::max<size_t>(need, cap * 2), (size_t)16 << 20);
        uint8_t *q = pinned ? (uint8_t *)tc_host_alloc(nc) : (uint8_t *)malloc(nc);
        if (q && !pinned && nc >= ((size_t)64 << 20)) {
            // first touch of fresh pages costs the single consumer thread ~0.7 s per GB: fault them in on all cores
            const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            vector<std::thread> th;
            for (unsigned t = 0; t < nt; t++)
                th.emplace_back([=]() { const size_t a = nc * t / nt, b = nc * (t + 1) / nt; for (size_t x = a; x < b; x += 4096) q[x] = 0; });
            for (std::thread &x : th) x.join();
        }
        if (!q) { cerr << "ERROR: cannot allocate " << nc << " bytes of " << (pinned ? "page-locked " : "") << "memory" << (pinned ? string(", ") + tc_last_error() : string()) << endl; exit(1); }
        if (len) memcpy(q, p, len);
        release();
        p = q; cap = nc;
    }
    void release() { if (p) { if (pinned) tc_host_free(p); else free(p); } p = nullptr; }
    void append(const uint8_t *src, size_t n) { reserve(len + n + 64); memcpy(p + len, src, n); len += n; }
    size_t size() const { return len; }
    ~PinnedBuf() { release(); }
};

struct Batch {
    PinnedBuf seq; vector<uint64_t> rec_off{0}; vector<uint32_t> rec2genome, ssize;
    vector<Genome> genomes;
    uint64_t longest = 0;
    void clear() { seq.len = 0; rec_off.assign(1, 0); rec2genome.clear(); ssize.clear(); genomes.clear(); longest = 0; }
};

void flush_batch(tc_ctx *ctx, const Options &o, Batch &b, vector<Genome> &out)
{
    if (b.genomes.empty()) return;
    const double t_flush = now_s();
    struct Report { double t0; size_t bytes, ng; ~Report() { if (getenv("TCLUST_TIMING")) cerr << "[tclust timing] sketch batch: " << ng << " genomes, " << bytes / 1e6 << " MB in " << (now_s() - t0) * 1e3 << " ms" << endl; } } report{t_flush, b.seq.len, b.genomes.size()};
    const uint32_t ng = (uint32_t)b.genomes.size();
    b.seq.reserve(b.seq.len + 64);
    memset(b.seq.p + b.seq.len, 0, 64);
    vector<uint64_t> hashes;
    vector<uint32_t> out_n(ng);
    uint64_t stride = 1;
    if (o.fast) {
        // survivors of the dimension filter: about L / 16^drlevel per genome; a repetitive genome can exceed any
        // estimate (the row holds them before deduplication), so the call is retried with a longer row
        stride = std::max<uint64_t>(1024, (b.longest >> (4 * o.drlevel)) * 2 + 1024);
        for (;;) {
            hashes.assign((size_t)ng * stride, 0);
            // list mode: sketchFileWithKssd; one multi-FASTA: the per-record consumer, which treats an odd k differently
            const int rc = (o.by_file ? tc_sketch_kssd : tc_sketch_kssd_seq)(ctx, b.seq.p, b.rec_off.data(), b.rec2genome.data(),
                                          b.rec2genome.size(), ng, o.kmer, o.drlevel, g_shuffled_dim.data(), hashes.data(), stride, out_n.data());
            if (!rc) break;
            if (rc == TC_ERR_KSSD_STRIDE && stride < b.longest + 1024) {      // grow to the reported need, not by a guess
                stride = std::min<uint64_t>(std::max<uint64_t>((uint64_t)tc_kssd_needed_stride() + 64, stride + 1024), b.longest + 1024);
                continue;
            }
            tc_or_die(rc, "kssd sketch generation");
        }
    } else {
        for (uint32_t s : b.ssize) stride = std::max<uint64_t>(stride, s);
        hashes.resize((size_t)ng * stride);
        tc_or_die(tc_sketch_minhash(ctx, b.seq.p, b.rec_off.data(), b.rec2genome.data(), b.rec2genome.size(), ng, o.kmer,
                                    b.ssize.data(), 42, hashes.data(), stride, out_n.data()), "sketch generation");
    }
    for (uint32_t g = 0; g < ng; g++) {
        b.genomes[g].hashes.assign(hashes.begin() + (size_t)g * stride, hashes.begin() + (size_t)g * stride + out_n[g]);
        out.push_back(std::move(b.genomes[g]));
    }
    b.clear();
}

// two batches: one is being filled by the parser while the other is sketched by a background thread
struct Feeder {
    tc_ctx *ctx; const Options &o; vector<Genome> &out;
    std::function<tc_ctx *()> late_ctx;       // set when the feeder was built before the context (pageable buffers)
    Batch batch[2]; int cur = 0; std::thread worker; bool running = false;
    Feeder(tc_ctx *c, const Options &opt, vector<Genome> &dst, uint64_t expect_bytes, uint64_t batch_bytes, bool pinned = true)
        : ctx(c), o(opt), out(dst)
    {
        batch[0].seq.pinned = batch[1].seq.pinned = pinned;
        // one allocation per buffer: the whole input if it fits a batch, else a full batch (+ the largest overshoot)
        const uint64_t first = std::min<uint64_t>(expect_bytes + (1u << 20), batch_bytes + (batch_bytes >> 3));
        batch[0].seq.reserve(first);
        // the second buffer is page-locked in the background while the first one fills
        if (expect_bytes > batch_bytes) second = std::thread([this, first]() { batch[1].seq.reserve(first); });
    }
    ~Feeder() { if (second.joinable()) second.join(); }
    std::thread second;
    Batch &fill() { return batch[cur]; }
    void wait() { if (running) { worker.join(); running = false; } }
    void submit()                      // hand the filled batch to the GPU, continue in the other one
    {
        wait();                        // results are appended in submission order
        if (second.joinable()) second.join();
        if (!ctx && late_ctx) ctx = late_ctx();
        Batch *b = &batch[cur];
        worker = std::thread([this, b]() { flush_batch(ctx, o, *b, out); });
        running = true;
        cur ^= 1;
    }
    void finish() { submit(); wait(); }
};

// list mode: one genome per file (sketchFiles)
void sketch_files(const std::function<tc_ctx *()> &get_ctx, const Options &o, vector<Genome> &out)
{
    fprintf(stderr, "-----input fileList, sketch by file\n");
    vector<string> files = read_list(o.input);
    const size_t nfiles = files.size();
    if (o.fast) {                                          // SketchInfo.cpp:1019-1031
        const int half_k = (o.kmer + 1) / 2;
        g_kssd.half_k = half_k;
        g_kssd.half_subk = 6 - o.drlevel >= 2 ? 6 : o.drlevel + 2;
        g_kssd.drlevel = o.drlevel;
        g_kssd.id = (half_k << 8) + (g_kssd.half_subk << 4) + o.drlevel;
        g_kssd.genome_number = (int)nfiles;
        make_shuffled_dim(g_kssd.half_subk);
    }
    const uint64_t batch_bytes = 1ull << 29;
    // Producer / consumer (RabbitFX's producer thread + kseq consumers, FastxStream.cpp:45-71,
    // SketchInfo.cpp:878-948): the host threads read and parse files ahead of the consumer -- up to `ahead` files in
    // flight -- while the consumer appends them, in list order, to the page-locked batch the GPU sketches next.  The
    // CUDA context (seconds to come up) is only waited for when the first batch buffer is allocated, so its start-up
    // overlaps reading and parsing.
    struct Parsed { vector<uint8_t> seq; vector<Record> recs; int file_len = 0; bool ok = false, ready = false; };
    vector<Parsed> parsed(nfiles);
    std::mutex mu;
    std::condition_variable cv_ready, cv_room;
    size_t consumed = 0, cursor = 0;
    // run-ahead of the parsers: about 2 GB of input in flight (average file size from cal_size), at least a few files per thread
    const size_t avg_file = std::max<uint64_t>(g_input_bytes / std::max<size_t>(nfiles, 1), 1);
    const size_t ahead = std::max<size_t>(std::max<size_t>((size_t)o.threads * 4, 16), (size_t)((2ull << 30) / avg_file));
    auto worker = [&]() {
        for (;;) {
            size_t i;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_room.wait(lk, [&]() { return cursor >= nfiles || cursor < consumed + ahead; });
                if (cursor >= nfiles) return;
                i = cursor++;
            }
            Parsed &p = parsed[i];
            static thread_local string text;                    // reused from file to file: no fresh pages per file
            if (read_whole_gz(files[i], text)) {
                parse_fasta(text, true, p.seq, p.recs);
                if (o.containment && !o.fast) { uint64_t u; file_size_like_reference(files[i], u, p.file_len); }
                p.ok = true;
            }
            { std::lock_guard<std::mutex> lk(mu); p.ready = true; }
            cv_ready.notify_all();
        }
    };
    vector<std::thread> pool;
    for (int t = 0; t < std::max(1, std::min<int>(o.threads, (int)nfiles)); t++) pool.emplace_back(worker);
    // The consumer only assigns offsets in the batch; a few copy threads move the parsed bytes there (one thread copies
    // into fresh page-locked memory at ~3 GB/s, which was the feeder's limit on inputs of several GB).  All copies of a
    // batch are drained before the batch is submitted or its buffer grows.
    struct CopyTask { uint8_t *dst; vector<uint8_t> src; };
    std::deque<CopyTask> copy_q;
    std::mutex cmu;
    std::condition_variable cv_copy, cv_drained;
    size_t copies_pending = 0;
    bool copy_stop = false;
    auto copier = [&]() {
        for (;;) {
            CopyTask t;
            {
                std::unique_lock<std::mutex> lk(cmu);
                cv_copy.wait(lk, [&]() { return copy_stop || !copy_q.empty(); });
                if (copy_q.empty()) return;
                t = std::move(copy_q.front());
                copy_q.pop_front();
            }
            memcpy(t.dst, t.src.data(), t.src.size());
            vector<uint8_t>().swap(t.src);
            { std::lock_guard<std::mutex> lk(cmu); copies_pending--; }
            cv_drained.notify_all();
        }
    };
    auto drain_copies = [&]() { std::unique_lock<std::mutex> lk(cmu); cv_drained.wait(lk, [&]() { return copies_pending == 0; }); };
    vector<std::thread> copiers;
    const int n_copiers = g_input_bytes > ((uint64_t)256 << 20) ? std::max(1, std::min(6, o.threads / 2)) : 0;
    for (int t = 0; t < n_copiers; t++) copiers.emplace_back(copier);
    std::unique_ptr<Feeder> feeder_p;
    const bool feeder_dry = getenv("TCLUST_FEEDER_DRY") != nullptr;
    const double t_feed0 = now_s();
    uint64_t dry_bytes = 0;
    double t_wait_parse = 0, t_wait_gpu = 0;
    for (size_t i = 0; i < nfiles; i++) {
        {
            const double tw = now_s();
            std::unique_lock<std::mutex> lk(mu);
            cv_ready.wait(lk, [&]() { return parsed[i].ready; });
            t_wait_parse += now_s() - tw;
        }
        Parsed &p = parsed[i];
        if (!p.ok) { fprintf(stderr, "cannot open the genome file: %s\n", files[i].c_str()); exit(1); }
        if (feeder_dry) {                                               // producer rate alone (tools/feeder_check.py --dry)
            dry_bytes += p.seq.size();
            vector<uint8_t>().swap(p.seq); vector<Record>().swap(p.recs);
            { std::lock_guard<std::mutex> lk(mu); consumed = i + 1; }
            cv_room.notify_all();
            if (i + 1 == nfiles) {
                cerr << "[tclust timing] feeder dry run: " << dry_bytes / 1e9 << " Gbp parsed in " << (now_s() - t_feed0) << " s with "
                     << pool.size() << " threads" << endl;
                for (std::thread &t : pool) t.join();
                exit(0);
            }
            continue;
        }
        uint64_t total_len = 0;
        for (const Record &r : p.recs) total_len += r.end - r.begin;
        if (total_len >= o.min_len) {                                  // SketchInfo.cpp:963
            if (!feeder_p) {
                // an input of at most two batches goes through plain memory and does not wait for the context before
                // its first batch is complete; larger inputs reuse two page-locked buffers
                const bool pinned = g_input_bytes > 2 * batch_bytes;
                const double tw = now_s();
                tc_ctx *c = pinned ? get_ctx() : nullptr;
                const double ta = now_s();
                feeder_p.reset(new Feeder(c, o, out, g_input_bytes, batch_bytes, pinned));
                feeder_p->late_ctx = get_ctx;
                if (getenv("TCLUST_TIMING") && pinned)
                    cerr << "[tclust timing] feeder: waited " << (ta - tw) * 1e3 << " ms for the context, page-locked batch buffers "
                         << (now_s() - ta) * 1e3 << " ms" << endl;
            }
            Feeder &feeder = *feeder_p;
            Genome g;
            g.file_name = files[i];
            g.name = "noName"; g.comment = "noName";                    // :934-939
            if (!p.recs.empty()) { g.name = p.recs[0].name; if (p.recs[0].has_comment) g.comment = p.recs[0].comment; }
            g.total_len = total_len;
            Batch &b = feeder.fill();
            const uint32_t gi = (uint32_t)b.genomes.size();
            const uint64_t base = b.seq.size();
            if (n_copiers && p.seq.size() >= ((size_t)256 << 10)) {
                if (b.seq.len + p.seq.size() + 64 > b.seq.cap) { drain_copies(); b.seq.reserve(b.seq.len + p.seq.size() + 64); }
                CopyTask t{b.seq.p + b.seq.len, std::move(p.seq)};
                b.seq.len += t.src.size();
                { std::lock_guard<std::mutex> lk(cmu); copy_q.push_back(std::move(t)); copies_pending++; }
                cv_copy.notify_one();
            } else {
                if (b.seq.len + p.seq.size() + 64 > b.seq.cap) drain_copies();
                b.seq.append(p.seq.data(), p.seq.size());
            }
            for (const Record &r : p.recs) { b.rec_off.push_back(base + r.end); b.rec2genome.push_back(gi); }
            b.ssize.push_back(o.containment ? (uint32_t)std::max(p.file_len / o.compress, 100) : (uint32_t)o.sketch_size);
            g.ssize = b.ssize.back();
            b.longest = std::max<uint64_t>(b.longest, total_len);
            b.genomes.push_back(std::move(g));
            if (b.seq.size() >= batch_bytes) { const double tw = now_s(); drain_copies(); feeder.submit(); t_wait_gpu += now_s() - tw; }
        }
        vector<uint8_t>().swap(p.seq); vector<Record>().swap(p.recs);
        { std::lock_guard<std::mutex> lk(mu); consumed = i + 1; }
        cv_room.notify_all();
    }
    for (std::thread &t : pool) t.join();
    drain_copies();
    { std::lock_guard<std::mutex> lk(cmu); copy_stop = true; }
    cv_copy.notify_all();
    for (std::thread &t : copiers) t.join();
    if (!feeder_p) { get_ctx(); return; }
    Feeder &feeder = *feeder_p;
    const double t_fin = now_s();
    feeder.finish();
    if (getenv("TCLUST_TIMING"))
        cerr << "[tclust timing] feeder: " << (now_s() - t_feed0) * 1e3 << " ms in all; the consumer waited " << t_wait_parse * 1e3
             << " ms for parsed files, " << t_wait_gpu * 1e3 << " ms for copies and the previous batch on the GPU, "
             << (now_s() - t_fin) * 1e3 << " ms for the last batch" << endl;
    for (size_t i = 0; i < out.size(); i++) out[i].id = (int)i;
}

// single-file mode: one genome per record (sketchSequences with the RabbitFX reader)
void sketch_sequences(tc_ctx *ctx, const Options &o, vector<Genome> &out)
{
    fprintf(stderr, "-----input one file, sketch by sequence: \n");
    if (o.fast) {
        // sketchSequencesWithKssd + consumer_fasta_task_with_kssd (src/SketchInfo.cpp:554-642, 274-433): one KSSD sketch per
        // record (tc_sketch_kssd_seq).  For an even k-mer size its window rule (`base > kmerSize`, no tuple reset on a
        // non-ACGT byte) selects exactly the windows of the list-mode loop; for an odd one the first window after every
        // gap keeps one stale base and the hash is packed with the un-rounded size -- K2 reproduces both.  (The
        // reference leaves the sets in hash-set iteration order; every consumer treats them as sets, we store them ascending.)
        const size_t dot = o.input.find_last_of('.');
        const string suf = dot == string::npos ? "" : o.input.substr(dot + 1);
        if (suf != "fasta" && suf != "fna" && suf != "fa") {           // SketchInfo.cpp:555-562
            cerr << "error input format file: " << o.input << endl << "Only support FASTA files" << endl;
            exit(0);
        }
        const int half_k = (o.kmer + 1) / 2;
        g_kssd.half_k = half_k;
        g_kssd.half_subk = 6 - o.drlevel >= 2 ? 6 : o.drlevel + 2;
        g_kssd.drlevel = o.drlevel;
        g_kssd.id = (half_k << 8) + (g_kssd.half_subk << 4) + o.drlevel;
        make_shuffled_dim(g_kssd.half_subk);
    }
    string text;
    if (!read_whole_gz(o.input, text)) die("cannot open the genome file, " + o.input);
    vector<uint8_t> seq; vector<Record> recs;
    parse_fasta(text, false, seq, recs);
    text.clear(); text.shrink_to_fit();
    Feeder feeder(ctx, o, out, seq.size(), 1ull << 29);
    for (const Record &r : recs) {
        const uint64_t len = r.end - r.begin;
        if (len < o.min_len) continue;                                   // SketchInfo.cpp:217
        Genome g;
        g.name = r.name; g.comment = r.comment; g.total_len = len;
        Batch &b = feeder.fill();
        const uint32_t gi = (uint32_t)b.genomes.size();
        b.seq.append(seq.data() + r.begin, len);
        b.rec_off.push_back(b.seq.size()); b.rec2genome.push_back(gi);
        b.ssize.push_back(o.containment ? (uint32_t)std::max((int)len / o.compress, 100) : (uint32_t)o.sketch_size);
        g.ssize = b.ssize.back();
        b.longest = std::max<uint64_t>(b.longest, len);
        b.genomes.push_back(std::move(g));
        if (b.seq.size() >= (1ull << 29)) feeder.submit();
    }
    feeder.finish();
    for (size_t i = 0; i < out.size(); i++) out[i].id = (int)i;
    if (o.fast) g_kssd.genome_number = (int)out.size();
}

// ---------------------------------------------------------------------------- on-disk formats
string current_date_time()
{
    time_t now = time(0);
    struct tm t = *localtime(&now);
    char buf[80];
    strftime(buf, sizeof buf, "%Y_%m_%d_%H-%M-%S", &t);
    return buf;
}

template <typename T> void wr(FILE *fp, const T &v) { fwrite(&v, sizeof(T), 1, fp); }
template <typename T> bool rd(FILE *fp, T &v) { return fread(&v, sizeof(T), 1, fp) == 1; }

void save_genome_info(const vector<Genome> &g, const string &folder, const char *type, bool by_file)
{
    string path = folder + "/info." + type;
    FILE *fp = fopen(path.c_str(), "w+");
    if (!fp) die("save_genome_info(), cannot open the file: " + path);
    wr(fp, by_file);
    wr(fp, (size_t)g.size());
    for (const Genome &x : g) {
        if (by_file) {
            wr(fp, (int)x.file_name.size()); wr(fp, (int)x.name.size()); wr(fp, (int)x.comment.size()); wr(fp, x.strand);
            wr(fp, (uint64_t)x.total_len);
            fwrite(x.file_name.data(), 1, x.file_name.size(), fp); fwrite(x.name.data(), 1, x.name.size(), fp);
            fwrite(x.comment.data(), 1, x.comment.size(), fp);
        } else {
            wr(fp, (int)x.name.size()); wr(fp, (int)x.comment.size()); wr(fp, x.strand); wr(fp, (int)x.total_len);
            fwrite(x.name.data(), 1, x.name.size(), fp); fwrite(x.comment.data(), 1, x.comment.size(), fp);
        }
    }
    fclose(fp);
}

bool load_genome_info(vector<Genome> &g, const string &folder, const char *type)
{
    string path = folder + "/info." + type;
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) die("load_genome_info(), cannot open the file: " + path);
    bool by_file = false; size_t n = 0;
    if (!rd(fp, by_file) || !rd(fp, n)) die("load_genome_info(), truncated file: " + path);
    g.resize(n);
    auto rs = [&](string &s, int len) { s.resize(len); if (len && fread(&s[0], 1, len, fp) != (size_t)len) die("truncated " + path); };
    for (size_t i = 0; i < n; i++) {
        Genome &x = g[i];
        x.id = (int)i;
        if (by_file) {
            int lf, ls, lc; uint64_t tot;
            if (!rd(fp, lf) || !rd(fp, ls) || !rd(fp, lc) || !rd(fp, x.strand) || !rd(fp, tot)) die("truncated " + path);
            x.total_len = tot; rs(x.file_name, lf); rs(x.name, ls); rs(x.comment, lc);
        } else {
            int ls, lc, len;
            if (!rd(fp, ls) || !rd(fp, lc) || !rd(fp, x.strand) || !rd(fp, len)) die("truncated " + path);
            x.total_len = (uint64_t)len; rs(x.name, ls); rs(x.comment, lc);
        }
    }
    fclose(fp);
    return by_file;
}

void save_sketches(const vector<Genome> &g, const string &folder, const Options &o)
{
    save_genome_info(g, folder, "sketch", o.by_file);
    string path = folder + "/hash.sketch";
    FILE *fp = fopen(path.c_str(), "w+");
    if (!fp) die("saveSketch(), cannot open the file: " + path);
    wr(fp, (int)0); wr(fp, o.kmer); wr(fp, o.containment); wr(fp, o.containment ? o.compress : o.sketch_size);
    for (const Genome &x : g) { wr(fp, (size_t)x.hashes.size()); fwrite(x.hashes.data(), 8, x.hashes.size(), fp); }
    fclose(fp);
    cerr << "-----save the sketches into: " << folder << endl;
}

bool load_sketches(vector<Genome> &g, const string &folder, Options &o)
{
    bool by_file = load_genome_info(g, folder, "sketch");
    string path = folder + "/hash.sketch";
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) die("loadSketches(), cannot open the file: " + path);
    int func_id = 0, param = 0;
    if (!rd(fp, func_id) || func_id != 0) die("loadSketches(), " + path + " does not hold MinHash sketches");
    if (!rd(fp, o.kmer) || !rd(fp, o.containment) || !rd(fp, param)) die("loadSketches(), truncated " + path);
    if (o.containment) o.compress = param; else o.sketch_size = param;
    for (Genome &x : g) {
        size_t n = 0;
        if (!rd(fp, n)) die("loadSketches(), truncated " + path);
        x.hashes.resize(n);
        if (n && fread(x.hashes.data(), 8, n, fp) != n) die("loadSketches(), truncated " + path);
        x.ssize = (uint32_t)param;               // new MinHash(kmer_size, contain_compress | sketch_size), Sketch_IO.cpp:333-340
    }
    fclose(fp);
    return by_file;
}

// ---- KSSD folder (src/Sketch_IO.cpp:36-86,136-167,228-282; index/dict src/SketchInfo.cpp:1379-1467)
bool kssd_use64(const KssdParams &k) { return k.half_k - k.drlevel > 8; }

void save_kssd_genome_info(const vector<Genome> &g, const string &folder, const char *type, bool by_file, bool use64)
{
    string path = folder + "/kssd.info." + type;
    FILE *fp = fopen(path.c_str(), "w+");
    if (!fp) die("save_genome_info(), cannot open the file: " + path);
    wr(fp, by_file);
    wr(fp, (size_t)g.size());
    for (const Genome &x : g) {
        if (by_file) {
            wr(fp, (int)x.file_name.size()); wr(fp, (int)x.name.size()); wr(fp, (int)x.comment.size()); wr(fp, x.strand);
            wr(fp, (uint64_t)x.total_len);
            fwrite(x.file_name.data(), 1, x.file_name.size(), fp); fwrite(x.name.data(), 1, x.name.size(), fp);
            fwrite(x.comment.data(), 1, x.comment.size(), fp);
        } else {
            wr(fp, (int)x.name.size()); wr(fp, (int)x.comment.size()); wr(fp, x.strand); wr(fp, (int)x.total_len);
            fwrite(x.name.data(), 1, x.name.size(), fp); fwrite(x.comment.data(), 1, x.comment.size(), fp);
        }
        wr(fp, use64);
    }
    fclose(fp);
}

bool load_kssd_genome_info(vector<Genome> &g, const string &folder, const char *type)
{
    string path = folder + "/kssd.info." + type;
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) die("load_kssd_genome_info(), cannot open the file: " + path);
    bool by_file = false; size_t n = 0;
    if (!rd(fp, by_file) || !rd(fp, n)) die("load_kssd_genome_info(), truncated file: " + path);
    g.resize(n);
    auto rs = [&](string &s, int len) { s.resize(len); if (len && fread(&s[0], 1, len, fp) != (size_t)len) die("truncated " + path); };
    for (size_t i = 0; i < n; i++) {
        Genome &x = g[i];
        x.id = (int)i;
        bool use64 = false;
        if (by_file) {
            int lf, ls, lc; uint64_t tot;
            if (!rd(fp, lf) || !rd(fp, ls) || !rd(fp, lc) || !rd(fp, x.strand) || !rd(fp, tot)) die("truncated " + path);
            x.total_len = tot; rs(x.file_name, lf); rs(x.name, ls); rs(x.comment, lc);
        } else {
            int ls, lc, len;
            if (!rd(fp, ls) || !rd(fp, lc) || !rd(fp, x.strand) || !rd(fp, len)) die("truncated " + path);
            x.total_len = (uint64_t)len; rs(x.name, ls); rs(x.comment, lc);
        }
        if (!rd(fp, use64)) die("truncated " + path);
    }
    fclose(fp);
    return by_file;
}

void save_kssd_sketches(const vector<Genome> &g, const string &folder, bool by_file)
{
    const bool use64 = kssd_use64(g_kssd);
    save_kssd_genome_info(g, folder, "sketch", by_file, use64);
    string path = folder + "/kssd.hash.sketch";
    FILE *fp = fopen(path.c_str(), "w+");
    if (!fp) die("saveSketch(), cannot open the file: " + path);
    wr(fp, g_kssd.id); wr(fp, g_kssd.half_k); wr(fp, g_kssd.half_subk); wr(fp, g_kssd.drlevel); wr(fp, g_kssd.genome_number);
    vector<uint32_t> tmp;
    for (const Genome &x : g) {
        wr(fp, (size_t)x.hashes.size());
        if (use64) fwrite(x.hashes.data(), 8, x.hashes.size(), fp);
        else { tmp.assign(x.hashes.begin(), x.hashes.end()); fwrite(tmp.data(), 4, tmp.size(), fp); }
    }
    fclose(fp);
    cerr << "-----save the kssd sketches into: " << folder << endl;
#ifdef GREEDY_CLUST
    return;     // clust-greedy writes no index files (sub_command.cpp:1943-1947)
#endif
    // kssd.sketch.index / kssd.sketch.dict: hash -> genome ids.  The reference's --fast --presketched reader has
    // no rebuild path and exits without them (src/MST.cpp:242-337), so a folder we write carries both.  Entries
    // are written in ascending hash order (the reference's order is its hash-map iteration order; the reader
    // only builds a map from them).
    vector<std::pair<uint64_t, uint32_t>> post;
    for (size_t i = 0; i < g.size(); i++) for (uint64_t h : g[i].hashes) post.emplace_back(h, (uint32_t)i);
    std::sort(post.begin(), post.end());
    vector<uint64_t> keys; vector<uint32_t> sizes, ids;
    for (size_t a = 0; a < post.size();) {
        size_t b = a;
        while (b < post.size() && post[b].first == post[a].first) ids.push_back(post[b++].second);
        keys.push_back(post[a].first); sizes.push_back((uint32_t)(b - a));
        a = b;
    }
    FILE *fd = fopen((folder + "/kssd.sketch.dict").c_str(), "w+");
    if (!fd) die("cannot open " + folder + "/kssd.sketch.dict");
    fwrite(ids.data(), 4, ids.size(), fd);
    fclose(fd);
    FILE *fi = fopen((folder + "/kssd.sketch.index").c_str(), "w+");
    if (!fi) die("cannot open " + folder + "/kssd.sketch.index");
    wr(fi, (size_t)keys.size());
    if (use64) fwrite(keys.data(), 8, keys.size(), fi);
    else { tmp.assign(keys.begin(), keys.end()); fwrite(tmp.data(), 4, tmp.size(), fi); }
    fwrite(sizes.data(), 4, sizes.size(), fi);
    fclose(fi);
}

bool load_kssd_sketches(vector<Genome> &g, const string &folder)
{
    string path = folder + "/kssd.hash.sketch";
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) {
        cerr << "ERROR: loadKssdSketches(), cannot open the file: " << path << endl;
        FILE *t = fopen((folder + "/hash.sketch").c_str(), "r");
        if (t) { cerr << "Do you want to load the minHash sketches directory? Try again without '--fast' option" << endl; fclose(t); }
        exit(1);
    }
    if (!rd(fp, g_kssd.id) || !rd(fp, g_kssd.half_k) || !rd(fp, g_kssd.half_subk) || !rd(fp, g_kssd.drlevel) ||
        !rd(fp, g_kssd.genome_number)) die("loadKssdSketches(), truncated " + path);
    const bool by_file = load_kssd_genome_info(g, folder, "sketch");
    const bool use64 = kssd_use64(g_kssd);
    vector<uint32_t> tmp;
    for (Genome &x : g) {
        size_t n = 0;
        if (!rd(fp, n)) die("loadKssdSketches(), truncated " + path);
        x.hashes.resize(n);
        if (use64) { if (n && fread(x.hashes.data(), 8, n, fp) != n) die("loadKssdSketches(), truncated " + path); }
        else {
            tmp.resize(n);
            if (n && fread(tmp.data(), 4, n, fp) != n) die("loadKssdSketches(), truncated " + path);
            std::copy(tmp.begin(), tmp.end(), x.hashes.begin());
        }
        // a folder the reference wrote from ONE multi-FASTA holds every sketch in hash-set iteration order
        // (consumer_fasta_task_with_kssd, SketchInfo.cpp:396-410); its consumers treat the arrays as sets
        if (!std::is_sorted(x.hashes.begin(), x.hashes.end())) std::sort(x.hashes.begin(), x.hashes.end());
    }
    fclose(fp);
    return by_file;
}

[[maybe_unused]] void save_mst(const vector<Genome> &g, const vector<tc_edge> &mst, const string &folder, bool by_file, bool fast)
{
    if (fast) save_kssd_genome_info(g, folder, "mst", by_file, kssd_use64(g_kssd));      // saveKssdMST, MST_IO.cpp:181-198
    else save_genome_info(g, folder, "mst", by_file);
    string path = folder + "/edge.mst";
    FILE *fp = fopen(path.c_str(), "w+");
    if (!fp) die("saveMST(), cannot open the file: " + path);
    wr(fp, (size_t)mst.size());
    for (const tc_edge &e : mst) { wr(fp, e.pre); wr(fp, e.suf); wr(fp, e.dist); }
    fclose(fp);
    cerr << (fast ? "-----save the kssd mst into: " : "-----save the mst into: ") << folder << endl;
}

[[maybe_unused]] bool load_mst(vector<Genome> &g, vector<tc_edge> &mst, const string &folder, bool fast)
{
    bool by_file = fast ? load_kssd_genome_info(g, folder, "mst") : load_genome_info(g, folder, "mst");
    string path = folder + "/edge.mst";
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) die("loadMST(), cannot open the file: " + path);
    size_t n = 0;
    if (!rd(fp, n)) die("loadMST(), truncated " + path);
    mst.resize(n);
    for (tc_edge &e : mst) if (!rd(fp, e.pre) || !rd(fp, e.suf) || !rd(fp, e.dist)) die("loadMST(), truncated " + path);
    fclose(fp);
    return by_file;
}

void print_result(const vector<vector<int>> &cluster, const vector<Genome> &g, bool by_file, const string &out, double threshold)
{
    FILE *fp = fopen(out.c_str(), "w");
    if (!fp) { cerr << "Error in printResult(), cannot open file: " << out << endl; exit(1); }
    if (threshold >= 0.0) {
        fprintf(fp, "# Clustering threshold: %.6f\n", threshold);
        fprintf(fp, "# Total clusters: %zu\n", cluster.size());
        fprintf(fp, "#\n");
    }
    for (size_t i = 0; i < cluster.size(); i++) {
        fprintf(fp, "the cluster %d is: \n", (int)i);
        for (size_t j = 0; j < cluster[i].size(); j++) {
            const int id = cluster[i][j];
            if (by_file)
                fprintf(fp, "\t%5d\t%6d\t%12dnt\t%20s\t%20s\t%s\n", (int)j, id, (int)g[id].total_len, g[id].file_name.c_str(),
                        g[id].name.c_str(), g[id].comment.c_str());
            else if (g[id].from_state)      // seqInfo of a state-loaded sketch is value-initialised: length 0, empty strings
                fprintf(fp, "\t%6d\t%6d\t%12dnt\t%20s\t%s\n", (int)j, id, 0, "", "");
            else
                fprintf(fp, "\t%6d\t%6d\t%12dnt\t%20s\t%s\n", (int)j, id, (int)g[id].total_len, g[id].name.c_str(),
                        g[id].comment.c_str());
        }
        fprintf(fp, "\n");
    }
    fclose(fp);
}

// ------------------------------------------------------------------------------- clustering
void upload(tc_ctx *ctx, const vector<Genome> &g)
{
    // one array per genome, as the reference holds them: the library gathers them through page-locked staging buffers
    vector<const uint64_t *> rows(g.size());
    vector<uint64_t> sizes(g.size());
    for (size_t i = 0; i < g.size(); i++) { rows[i] = g[i].hashes.data(); sizes[i] = g[i].hashes.size(); }
    tc_or_die(tc_sketches_upload_rows(ctx, rows.data(), sizes.data(), (uint32_t)g.size()), "sketch upload");
}

[[maybe_unused]] vector<vector<int>> clusters_from_forest(const vector<tc_edge> &mst, double threshold, int vertices)
{
    vector<vector<int>> adj(vertices);
    for (const tc_edge &e : mst)
        if (e.dist <= threshold) { adj[e.pre].push_back(e.suf); adj[e.suf].push_back(e.pre); }   // generateForest
    vector<char> visited(vertices, 0);
    vector<vector<int>> res;
    for (int i = 0; i < vertices; i++) {                                                        // generateClusterWithBfs
        if (visited[i]) continue;
        visited[i] = 1;
        std::queue<int> q;
        q.push(i);
        vector<int> c{i};
        while (!q.empty()) {
            int k = q.front(); q.pop();
            for (int v : adj[k]) if (!visited[v]) { visited[v] = 1; q.push(v); c.push_back(v); }
        }
        res.push_back(std::move(c));
    }
    return res;
}

// ---- MST consumers: single-linkage dendrogram as Newick / PHYLIP / NEXUS text and as a linkage matrix
// (get_newick_tree src/MST.cpp:1088-1147, get_linkage_from_mst :1235-1284, printers src/MST_IO.cpp:263-375).
// Edges are merged in ascending distance (std::sort on dist alone, as the reference); every merge creates an inner
// node at height = edge distance whose children hang on branches of length max(0, height - child height); the tree
// printed is the component of genome 0.
struct MergeForest {                      // union-find by rank; find() returns the set's root element
    vector<int> p, r;
    explicit MergeForest(int n) : p(n), r(n, 0) { for (int i = 0; i < n; i++) p[i] = i; }
    int find(int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    int unite(int a, int b)
    {
        a = find(a); b = find(b);
        if (a == b) return a;
        if (r[a] < r[b]) std::swap(a, b);
        p[b] = a;
        if (r[a] == r[b]) r[a]++;
        return a;
    }
};

vector<tc_edge> edges_by_distance(const vector<tc_edge> &mst)
{
    vector<tc_edge> e = mst;
    std::sort(e.begin(), e.end(), [](const tc_edge &a, const tc_edge &b) { return a.dist < b.dist; });
    return e;
}

string newick_of(const vector<Genome> &g, const vector<tc_edge> &mst, bool by_file)
{
    const int n = (int)g.size();
    auto label = [&](int i) { return by_file ? g[i].file_name : g[i].name; };
    if (n == 0) return ";";
    if (n == 1) return label(0) + ";";
    struct Child { int node; double len; };
    vector<std::pair<Child, Child>> kids(2 * n - 1, {{-1, 0.0}, {-1, 0.0}});
    vector<double> height(2 * n - 1, 0.0);
    vector<int> top(n);                                   // dendrogram node on top of each set root
    for (int i = 0; i < n; i++) top[i] = i;
    MergeForest uf(n);
    int next = n;
    for (const tc_edge &e : edges_by_distance(mst)) {
        const int ru = uf.find(e.pre), rv = uf.find(e.suf);
        if (ru == rv) continue;
        const int a = top[ru], b = top[rv];
        kids[next] = {{a, std::max(0.0, e.dist - height[a])}, {b, std::max(0.0, e.dist - height[b])}};
        height[next] = e.dist;
        top[uf.unite(ru, rv)] = next++;
    }
    // iterative pre-order emission (a 100k-leaf caterpillar tree would overflow a recursive walk)
    string out;
    struct Frame { int node; int stage; double len; bool has_len; };
    vector<Frame> st;
    st.push_back({top[uf.find(0)], 0, 0.0, false});
    while (!st.empty()) {
        Frame &f = st.back();
        const int node = f.node;
        if (kids[node].first.node < 0) {                   // leaf
            out += label(node);
            if (f.has_len) out += ":" + std::to_string(f.len);
            st.pop_back();
            continue;
        }
        if (f.stage == 0) { out += "("; f.stage = 1; st.push_back({kids[node].first.node, 0, kids[node].first.len, true}); }
        else if (f.stage == 1) { out += ","; f.stage = 2; st.push_back({kids[node].second.node, 0, kids[node].second.len, true}); }
        else {
            out += ")";
            if (f.has_len) out += ":" + std::to_string(f.len);
            st.pop_back();
        }
    }
    return out + ";";
}

[[maybe_unused]] void write_tree_outputs(const vector<Genome> &g, const vector<tc_edge> &mst, bool by_file, const Options &o)
{
    auto open_or_die = [](const string &path) { FILE *fp = fopen(path.c_str(), "w"); if (!fp) die("cannot write file: " + path); return fp; };
    string tree;
    if (o.newick || o.phylip || o.nexus) tree = newick_of(g, mst, by_file);
    if (o.newick) {
        const string path = o.output + ".newick.tree";
        FILE *fp = open_or_die(path); fprintf(fp, "%s\n", tree.c_str()); fclose(fp);
        cerr << "-----write the newick tree into: " << path << endl;
    }
    if (o.phylip) {
        const string path = o.output + ".phylip.tree";
        FILE *fp = open_or_die(path); fprintf(fp, "1\n%s\n", tree.c_str()); fclose(fp);
        cerr << "-----write the PHYLIP tree into: " << path << endl;
    }
    if (o.nexus) {
        const string path = o.output + ".nexus.tree";
        FILE *fp = open_or_die(path);
        fprintf(fp, "#NEXUS\nBEGIN TAXA;\n  DIMENSIONS NTAX=%zu;\n  TAXLABELS", g.size());
        for (const Genome &x : g) {
            string lab = by_file ? x.file_name : x.name;
            for (size_t pos = 0; (pos = lab.find('\'', pos)) != string::npos; pos += 2) lab.insert(pos, "'");
            fprintf(fp, " '%s'", lab.c_str());
        }
        fprintf(fp, ";\nEND;\nBEGIN TREES;\n  TREE tree_1 = [&R] %s\nEND;\n", tree.c_str());
        fclose(fp);
        cerr << "-----write the NEXUS tree into: " << path << endl;
    }
    if (o.linkage) {
        const string path = o.output + ".linkage.txt";
        FILE *fp = open_or_die(path);
        const int n = (int)g.size();
        if (n > 1) {
            MergeForest uf(n);
            vector<int> cid(n), csize(2 * n - 1, 1);
            for (int i = 0; i < n; i++) cid[i] = i;
            int next = n;
            for (const tc_edge &e : edges_by_distance(mst)) {
                const int ru = uf.find(e.pre), rv = uf.find(e.suf);
                if (ru == rv) continue;
                const int a = cid[ru], b = cid[rv];
                csize[next] = csize[a] + csize[b];
                fprintf(fp, "%d\t%d\t%.6f\t%d\n", a, b, e.dist, csize[next]);
                cid[uf.unite(ru, rv)] = next++;
            }
        }
        fclose(fp);
        cerr << "-----write the linkage matrix into: " << path << endl;
    }
}

#ifdef GREEDY_CLUST
// ------------------------------------------------------------------ clust-greedy --fast --append (cluster state)
// KssdClusterState (src/greedy.h:47-72) and its file cluster_state.bin (KssdClusterState::save / load,
// src/greedy.cpp:1545-1734): threshold, kmer_size, half_k, half_subk, drlevel, genomeNumber; representative ids; every
// sketch {int id, u64 totalSeqLength, bool use64, u32 sketchsize, size_t n32, size_t n64, hashes, size_t nameLen, fileName};
// the clusters; "KSSI02\0\0" + the hash -> representative-number index.  The reference writes the index in the iteration
// order of its hash map (it differs from run to run); this writer emits ascending hashes, the reader builds a map either way.
struct KssdState {
    double threshold = 0.0; int kmer_size = 0; KssdParams params;
    vector<int> rep_ids; vector<Genome> all; vector<vector<int>> clusters;
};

bool save_kssd_state(const KssdState &s, const string &path)
{
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) { cerr << "ERROR: Cannot open file for writing: " << path << endl; return false; }
    const bool use64 = kssd_use64(s.params);
    wr(fp, s.threshold); wr(fp, s.kmer_size);
    wr(fp, s.params.half_k); wr(fp, s.params.half_subk); wr(fp, s.params.drlevel); wr(fp, s.params.genome_number);
    wr(fp, (size_t)s.rep_ids.size());
    fwrite(s.rep_ids.data(), sizeof(int), s.rep_ids.size(), fp);
    wr(fp, (size_t)s.all.size());
    vector<uint32_t> tmp;
    for (const Genome &x : s.all) {
        wr(fp, x.id); wr(fp, (uint64_t)x.total_len); wr(fp, use64); wr(fp, (uint32_t)x.hashes.size());
        wr(fp, use64 ? (size_t)0 : (size_t)x.hashes.size());
        wr(fp, use64 ? (size_t)x.hashes.size() : (size_t)0);
        if (use64) fwrite(x.hashes.data(), 8, x.hashes.size(), fp);
        else { tmp.assign(x.hashes.begin(), x.hashes.end()); fwrite(tmp.data(), 4, tmp.size(), fp); }
        wr(fp, (size_t)x.file_name.size());
        fwrite(x.file_name.data(), 1, x.file_name.size(), fp);
    }
    wr(fp, (size_t)s.clusters.size());
    for (const vector<int> &c : s.clusters) { wr(fp, (size_t)c.size()); fwrite(c.data(), sizeof(int), c.size(), fp); }
    vector<std::pair<uint64_t, int>> post;
    for (size_t r = 0; r < s.rep_ids.size(); r++) for (uint64_t h : s.all[s.rep_ids[r]].hashes) post.emplace_back(h, (int)r);
    std::sort(post.begin(), post.end());
    size_t distinct = 0;
    for (size_t a = 0; a < post.size(); a++) distinct += a == 0 || post[a].first != post[a - 1].first;
    const char magic[8] = {'K', 'S', 'S', 'I', '0', '2', '\0', '\0'};
    fwrite(magic, 1, 8, fp);
    wr(fp, distinct);
    cerr << "Saving inverted index: " << distinct << " unique hashes..." << endl;
    for (size_t a = 0; a < post.size();) {
        size_t b = a;
        while (b < post.size() && post[b].first == post[a].first) b++;
        wr(fp, post[a].first); wr(fp, (size_t)(b - a));
        for (size_t e = a; e < b; e++) wr(fp, post[e].second);
        a = b;
    }
    const bool ok = !ferror(fp);
    fclose(fp);
    cerr << "Saved clustering state to: " << path << endl;
    cerr << "  - " << s.all.size() << " genomes" << endl;
    cerr << "  - " << s.rep_ids.size() << " clusters (representatives)" << endl;
    cerr << "  - " << distinct << " unique hashes in inverted index" << endl;
    return ok;
}

bool load_kssd_state(KssdState &s, const string &path)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) { cerr << "ERROR: Cannot open file for reading: " << path << endl; return false; }
    auto bad = [&]() { fclose(fp); cerr << "ERROR: " << path << " is truncated or not a KSSD cluster state" << endl; return false; };
    if (!rd(fp, s.threshold) || !rd(fp, s.kmer_size) || !rd(fp, s.params.half_k) || !rd(fp, s.params.half_subk) ||
        !rd(fp, s.params.drlevel) || !rd(fp, s.params.genome_number)) return bad();
    s.params.id = (s.params.half_k << 8) + (s.params.half_subk << 4) + s.params.drlevel;
    size_t nrep = 0, nsk = 0, ncl = 0;
    if (!rd(fp, nrep) || nrep > ((size_t)1 << 31)) return bad();
    s.rep_ids.resize(nrep);
    if (nrep && fread(s.rep_ids.data(), sizeof(int), nrep, fp) != nrep) return bad();
    if (!rd(fp, nsk) || nsk > ((size_t)1 << 31)) return bad();
    s.all.assign(nsk, Genome());
    vector<uint32_t> tmp;
    for (Genome &x : s.all) {
        bool use64 = false; uint32_t sketchsize = 0; uint64_t len = 0; size_t n32 = 0, n64 = 0, nl = 0;
        if (!rd(fp, x.id) || !rd(fp, len) || !rd(fp, use64) || !rd(fp, sketchsize) || !rd(fp, n32) || !rd(fp, n64)) return bad();
        if (n32 > ((size_t)1 << 32) || n64 > ((size_t)1 << 32) || (n32 && n64)) return bad();
        x.total_len = len;
        if (n32) {
            tmp.resize(n32);
            if (fread(tmp.data(), 4, n32, fp) != n32) return bad();
            x.hashes.assign(tmp.begin(), tmp.end());
        }
        if (n64) { x.hashes.resize(n64); if (fread(x.hashes.data(), 8, n64, fp) != n64) return bad(); }
        if (!std::is_sorted(x.hashes.begin(), x.hashes.end())) std::sort(x.hashes.begin(), x.hashes.end());
        if (!rd(fp, nl) || nl > ((size_t)1 << 20)) return bad();
        x.file_name.resize(nl);
        if (nl && fread(&x.file_name[0], 1, nl, fp) != nl) return bad();
        x.name = "N/A"; x.comment = "N/A";           // the state keeps no record names: printKssdResult prints N/A (MST_IO.cpp:99-104)
        x.from_state = true;
    }
    for (int r : s.rep_ids) if (r < 0 || (size_t)r >= nsk) return bad();
    if (!rd(fp, ncl) || ncl > nsk) return bad();
    s.clusters.assign(ncl, vector<int>());
    for (vector<int> &c : s.clusters) {
        size_t m = 0;
        if (!rd(fp, m) || m > nsk) return bad();
        c.resize(m);
        if (m && fread(c.data(), sizeof(int), m, fp) != m) return bad();
    }
    // the hash -> representative index that follows is rebuilt on the device from the representatives' sketches
    fclose(fp);
    if (ncl != nrep) { cerr << "ERROR: " << path << " holds " << ncl << " clusters for " << nrep << " representatives" << endl; return false; }
    cerr << "Loaded clustering state from: " << path << endl;
    cerr << "  - " << nsk << " genomes" << endl;
    cerr << "  - " << nrep << " clusters (representatives)" << endl;
    return true;
}

// MinHashClusterState (src/greedy.h:84-112) and its cluster_state.bin (MinHashClusterState::save / load,
// src/greedy.cpp:2134-2290): "MINHASH\0", threshold, kmer_size, sketch_size, bool is_containment; representative ids; every
// sketch {int id, u64 totalSeqLength, size_t n, u64 hashes[n], size_t nameLen, fileName} (the loader skips them: the
// sketches are reloaded from the folder); the clusters; the hash -> representative-number index (no marker here).
struct MinHashState {
    double threshold = 0.0; int kmer_size = 0, sketch_size = 0; bool containment = false;
    vector<int> rep_ids; vector<vector<int>> clusters;
};

bool save_minhash_state(const MinHashState &s, const vector<Genome> &all, const string &path)
{
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) { cerr << "ERROR: Cannot open file for writing: " << path << endl; return false; }
    const char magic[8] = {'M', 'I', 'N', 'H', 'A', 'S', 'H', '\0'};
    fwrite(magic, 1, 8, fp);
    wr(fp, s.threshold); wr(fp, s.kmer_size); wr(fp, s.sketch_size); wr(fp, s.containment);
    wr(fp, (size_t)s.rep_ids.size());
    fwrite(s.rep_ids.data(), sizeof(int), s.rep_ids.size(), fp);
    wr(fp, (size_t)all.size());
    for (const Genome &x : all) {
        wr(fp, x.id); wr(fp, (uint64_t)x.total_len);
        wr(fp, (size_t)x.hashes.size());
        fwrite(x.hashes.data(), 8, x.hashes.size(), fp);
        wr(fp, (size_t)x.file_name.size());
        fwrite(x.file_name.data(), 1, x.file_name.size(), fp);
    }
    wr(fp, (size_t)s.clusters.size());
    for (const vector<int> &c : s.clusters) { wr(fp, (size_t)c.size()); fwrite(c.data(), sizeof(int), c.size(), fp); }
    vector<std::pair<uint64_t, int>> post;
    for (size_t r = 0; r < s.rep_ids.size(); r++) for (uint64_t h : all[s.rep_ids[r]].hashes) post.emplace_back(h, (int)r);
    std::sort(post.begin(), post.end());
    size_t distinct = 0;
    for (size_t a = 0; a < post.size(); a++) distinct += a == 0 || post[a].first != post[a - 1].first;
    wr(fp, distinct);
    cerr << "Saving inverted index: " << distinct << " unique hashes..." << endl;
    for (size_t a = 0; a < post.size();) {
        size_t b = a;
        while (b < post.size() && post[b].first == post[a].first) b++;
        wr(fp, post[a].first); wr(fp, (size_t)(b - a));
        for (size_t e = a; e < b; e++) wr(fp, post[e].second);
        a = b;
    }
    const bool ok = !ferror(fp);
    fclose(fp);
    cerr << "Saved clustering state to: " << path << endl;
    cerr << "  - " << all.size() << " genomes" << endl;
    cerr << "  - " << s.rep_ids.size() << " clusters (representatives)" << endl;
    cerr << "  - " << distinct << " unique hashes in inverted index" << endl;
    return ok;
}

bool load_minhash_state(MinHashState &s, const string &path)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) { cerr << "ERROR: Cannot open file for reading: " << path << endl; return false; }
    auto bad = [&]() { fclose(fp); cerr << "ERROR: " << path << " is truncated or not a MinHash cluster state" << endl; return false; };
    char magic[8];
    if (fread(magic, 1, 8, fp) != 8 || strncmp(magic, "MINHASH", 8) != 0) {
        fclose(fp);
        cerr << "ERROR: Invalid file format (not a MinHash cluster state)" << endl;
        return false;
    }
    if (!rd(fp, s.threshold) || !rd(fp, s.kmer_size) || !rd(fp, s.sketch_size) || !rd(fp, s.containment)) return bad();
    size_t nrep = 0, nsk = 0, ncl = 0;
    if (!rd(fp, nrep) || nrep > ((size_t)1 << 31)) return bad();
    s.rep_ids.resize(nrep);
    if (nrep && fread(s.rep_ids.data(), sizeof(int), nrep, fp) != nrep) return bad();
    if (!rd(fp, nsk) || nsk > ((size_t)1 << 31)) return bad();
    for (size_t i = 0; i < nsk; i++) {                     // skipped: the sketches are reloaded from the folder (greedy.cpp:2232-2250)
        int id; uint64_t len; size_t nh = 0, nl = 0;
        if (!rd(fp, id) || !rd(fp, len) || !rd(fp, nh) || nh > ((size_t)1 << 32) || fseek(fp, (long)(nh * 8), SEEK_CUR) != 0) return bad();
        if (!rd(fp, nl) || nl > ((size_t)1 << 20) || fseek(fp, (long)nl, SEEK_CUR) != 0) return bad();
    }
    if (!rd(fp, ncl) || ncl > ((size_t)1 << 31)) return bad();
    s.clusters.assign(ncl, vector<int>());
    for (vector<int> &c : s.clusters) {
        size_t m = 0;
        if (!rd(fp, m) || m > ((size_t)1 << 31)) return bad();
        c.resize(m);
        if (m && fread(c.data(), sizeof(int), m, fp) != m) return bad();
    }
    fclose(fp);
    cerr << "Loaded clustering state from: " << path << endl;
    cerr << "  - " << nsk << " genomes" << endl;
    cerr << "  - " << nrep << " clusters (representatives)" << endl;
    return true;
}

// append_clust_greedy with a saved state (src/sub_command.cpp:97-190): true when the state was found and used
bool append_greedy_minhash_state(tc_ctx *ctx, Options &o)
{
    const string state_file = o.folder + "/cluster_state.bin";
    struct stat sb;
    if (stat(state_file.c_str(), &sb) != 0) return false;
    cerr << "===== Incremental Update Mode (MinHash) =====" << endl;
    cerr << "Found existing cluster state, loading..." << endl;
    MinHashState st;
    if (!load_minhash_state(st, state_file)) return false;          // the reference falls back to the classic append
    cerr << "---the threshold is: " << o.threshold << endl;
    cerr << "---the thread number is: " << o.threads << endl;
    // the sketches come from the folder again, in the folder's order (:101-104); the representatives must lie in it
    vector<Genome> all;
    const bool pre_by_file = load_sketches(all, o.folder, o);        // also sets o.kmer / o.containment / o.compress / o.sketch_size
    if (pre_by_file != o.by_file) cerr << "Warning: the input format of append genomes and pre-sketched genome is not same" << endl;
    for (int r : st.rep_ids)
        if (r < 0 || (size_t)r >= all.size()) {
            cerr << "ERROR: Representative ID " << r << " out of range [0, " << all.size() << "), cannot continue" << endl;
            exit(1);
        }
    cerr << "Successfully loaded " << st.rep_ids.size() << " representatives" << endl;
    if (st.kmer_size <= 0 || st.sketch_size <= 0) {
        cerr << "ERROR: Invalid kmer_size or sketch_size: kmer=" << st.kmer_size << ", sketch=" << st.sketch_size << endl;
        exit(1);
    }
    if (st.containment && o.compress <= 0) { cerr << "ERROR: is_containment is true but contain_compress is " << o.compress << endl; exit(1); }
    if (st.clusters.size() != st.rep_ids.size()) die(state_file + " holds " + std::to_string(st.clusters.size()) + " clusters for " +
                                                     std::to_string(st.rep_ids.size()) + " representatives");
    cerr << "Computing sketches for new genomes..." << endl;
    cerr << "  Parameters: kmer_size=" << st.kmer_size << ", sketch_size=" << st.sketch_size << ", is_containment=" << st.containment
         << ", contain_compress=" << o.compress << endl;
    o.kmer = st.kmer_size; o.sketch_size = st.sketch_size; o.containment = st.containment;
    vector<Genome> add;
    if (o.by_file) sketch_files([&]() { return ctx; }, o, add); else sketch_sequences(ctx, o, add);
    cerr << "New genomes sketched: " << add.size() << endl;
    if (add.empty()) { cerr << "ERROR: No new sketches generated" << endl; return true; }
    cerr << "Starting incremental clustering..." << endl;
    const uint32_t n_old = (uint32_t)all.size();
    for (Genome &x : add) all.push_back(std::move(x));
    const uint32_t n = (uint32_t)all.size();
    upload(ctx, all);
    vector<uint32_t> seeds(st.rep_ids.begin(), st.rep_ids.end());
    vector<int32_t> rep_of(n, -1);
    uint32_t nclust = 0;
    tc_or_die(tc_greedy_incremental(ctx, st.kmer_size, st.containment, st.threshold, n_old, seeds.data(), (uint32_t)seeds.size(),
                                    rep_of.data(), &nclust), "incremental greedy clustering");
    vector<int> cid(n, -1);
    for (size_t r = 0; r < st.rep_ids.size(); r++) cid[st.rep_ids[r]] = (int)r;
    int n_new = 0, n_assigned = 0;
    for (uint32_t q = n_old; q < n; q++) {
        if (rep_of[q] == (int32_t)q) {          // opens an EMPTY cluster, as in the KSSD flow (greedy.cpp:2099-2104)
            cid[q] = (int)st.clusters.size();
            st.rep_ids.push_back((int)q);
            st.clusters.emplace_back();
            n_new++;
        } else {
            st.clusters[cid[rep_of[q]]].push_back((int)q);
            n_assigned++;
        }
    }
    cerr << "\n===== Incremental Clustering Results =====" << endl;
    cerr << "Assigned to existing clusters: " << n_assigned << endl;
    cerr << "New clusters created: " << n_new << endl;
    cerr << "Total clusters now: " << st.clusters.size() << endl;
    cerr << "Incremental clustering completed" << endl;
    if (!o.no_save && o.save_rep) save_minhash_state(st, all, state_file);
    print_result(st.clusters, all, o.by_file, o.output, -1.0);
    cerr << "-----write the cluster result into: " << o.output << endl;
    cerr << "-----the cluster number of " << o.output << " is: " << st.clusters.size() << endl;
    cerr << "-----updated cluster state saved" << endl;
    return true;
}

// append_clust_greedy_fast (src/sub_command.cpp:192-276)
int append_greedy_fast(tc_ctx *ctx, Options &o)
{
    const string state_file = o.folder + "/cluster_state.bin";
    const bool save = !o.no_save;
    KssdState st;
    struct stat sb;
    bool has_state = stat(state_file.c_str(), &sb) == 0;
    if (has_state) {
        cerr << "===== Incremental Update Mode (KSSD) =====" << endl;
        cerr << "Found existing cluster state, loading..." << endl;
        has_state = load_kssd_state(st, state_file);
    }
    auto sketch_new = [&](vector<Genome> &add, const KssdParams &want) {
        double t0 = now_s();
        if (o.by_file) sketch_files([&]() { return ctx; }, o, add); else sketch_sequences(ctx, o, add);
        if (g_kssd.half_k != want.half_k || g_kssd.half_subk != want.half_subk || g_kssd.drlevel != want.drlevel)
            die("append_clust_greedy_fast(), the KSSD parameters of " + o.folder + " cannot be reproduced for the new genomes");
        cerr << "-----the size of sketches (number of genomes or sequences) is: " << add.size() << endl;
        cerr << "========time of computing sketch is: " << now_s() - t0 << "========" << endl;
        if (save) {                                              // compute_kssd_sketches: the new sketches get a folder of their own
            const string folder = current_date_time();
            if (system(("mkdir -p " + folder).c_str()) != 0) die("cannot create " + folder);
            save_kssd_sketches(add, folder, o.by_file);
        }
    };
    vector<Genome> add;
    if (!has_state) {
        cerr << "===== Initial State Building Mode (KSSD) =====" << endl;
        cerr << "No existing state found, building state from pre-sketched genomes..." << endl;
        vector<Genome> pre;
        const bool pre_by_file = load_kssd_sketches(pre, o.folder);
        const KssdParams pre_info = g_kssd;
        if (pre_by_file != o.by_file) cerr << "Warning: the input format of append genomes and pre-sketched genome is not same" << endl;
        if (pre.empty()) die("no sketches in " + o.folder);
        o.kmer = 2 * pre_info.half_k;
        o.drlevel = pre_info.drlevel;
        cerr << "-----use the same sketch parameters with pre-generated sketches" << endl;
        cerr << "---use the KSSD sketches" << endl;
        cerr << "---the half_k is: " << pre_info.half_k << endl;
        cerr << "---the half_subk is: " << pre_info.half_subk << endl;
        cerr << "---the drlevel is: " << o.drlevel << endl;
        cerr << "---the threshold is: " << o.threshold << endl;
        sketch_new(add, pre_info);
        g_kssd = pre_info;
        // KssdInitialClusterWithState (greedy.cpp:902-957): the greedy pass of clust-greedy --fast, its sorted sketch
        // order kept as the order of the state
        std::sort(pre.begin(), pre.end(), [](const Genome &a, const Genome &b) { return a.hashes.size() > b.hashes.size(); });
        upload(ctx, pre);
        vector<int32_t> rep_of(pre.size());
        uint32_t nclust = 0;
        tc_or_die(tc_greedy_cluster_kssd(ctx, o.kmer, o.threshold, rep_of.data(), &nclust), "greedy clustering");
        vector<int> cid(pre.size(), -1);
        for (int j = 0; j < (int)pre.size(); j++) {
            if (rep_of[j] == j) { cid[j] = (int)st.clusters.size(); st.clusters.push_back({j}); st.rep_ids.push_back(j); }
            else st.clusters[cid[rep_of[j]]].push_back(j);
        }
        st.all = std::move(pre);
        st.params = pre_info; st.threshold = o.threshold; st.kmer_size = o.kmer;
    } else {
        cerr << "---the threshold is: " << o.threshold << endl;
        cerr << "---the thread number is: " << o.threads << endl;
        o.kmer = st.kmer_size;
        o.drlevel = st.params.drlevel;
        sketch_new(add, st.params);
        g_kssd = st.params;
        cerr << "New genomes sketched: " << add.size() << endl;
    }
    // KssdIncrementalCluster (greedy.cpp:1736-1880) with the state's own threshold and k-mer size
    const uint32_t n_old = (uint32_t)st.all.size();
    for (Genome &x : add) st.all.push_back(std::move(x));
    const uint32_t n = (uint32_t)st.all.size();
    upload(ctx, st.all);
    vector<uint32_t> seeds(st.rep_ids.begin(), st.rep_ids.end());
    vector<int32_t> rep_of(n, -1);
    uint32_t nclust = 0;
    tc_or_die(tc_greedy_incremental_kssd(ctx, st.kmer_size, st.threshold, n_old, seeds.data(), (uint32_t)seeds.size(), rep_of.data(), &nclust),
              "incremental greedy clustering");
    vector<int> cid(n, -1);
    for (size_t r = 0; r < st.rep_ids.size(); r++) cid[st.rep_ids[r]] = (int)r;
    int n_new = 0, n_assigned = 0;
    for (uint32_t q = n_old; q < n; q++) {
        if (rep_of[q] == (int32_t)q) {
            // a new representative opens an EMPTY cluster: the reference does not list it among its own members
            // (greedy.cpp:1851-1856), so it never appears in the cluster file
            cid[q] = (int)st.clusters.size();
            st.rep_ids.push_back((int)q);
            st.clusters.emplace_back();
            n_new++;
        } else {
            st.clusters[cid[rep_of[q]]].push_back((int)q);
            n_assigned++;
        }
    }
    cerr << "\n===== Incremental Clustering Results =====" << endl;
    cerr << "Assigned to existing clusters: " << n_assigned << endl;
    cerr << "New clusters created: " << n_new << endl;
    cerr << "Total clusters now: " << st.clusters.size() << endl;
    if (save && o.save_rep) save_kssd_state(st, state_file);
    print_result(st.clusters, st.all, o.by_file, o.output, -1.0);
    cerr << "-----write the cluster result into: " << o.output << endl;
    cerr << "-----the cluster number of " << o.output << " is: " << st.clusters.size() << endl;
    cerr << (has_state ? "-----updated cluster state saved" : "-----saved cluster state (with inverted index) for future incremental updates") << endl;
    return 0;
}
#endif

void run_clustering(tc_ctx *ctx, vector<Genome> &g, const Options &o, bool by_file, bool save, const string &folder)
{
    if (g.empty()) die("no sketches to cluster");
    double t2 = now_s();
#ifdef GREEDY_CLUST
    // KssdGreedyClusterWithInvertedIndex (greedy.cpp:592-599): std::sort by hash count, descending.  The sort is not
    // stable; the same libstdc++ algorithm on the same key sequence yields the same permutation, which is what
    // makes the cluster file comparable with the reference's.
    if (o.fast) std::sort(g.begin(), g.end(), [](const Genome &a, const Genome &b) { return a.hashes.size() > b.hashes.size(); });
#endif
    upload(ctx, g);
    if (getenv("TCLUST_TIMING")) cerr << "[tclust timing] upload of the sketches " << (now_s() - t2) * 1e3 << " ms" << endl;
#ifdef GREEDY_CLUST
    vector<int32_t> rep_of(g.size());
    uint32_t nclust = 0;
    if (o.gpus > 1) {
        // representative scan sharded over the devices: every device holds all sketches, the representative
        // table is dealt round-robin; same clusters as with one GPU
        const int G = o.gpus;
        vector<tc_ctx *> cs(G, nullptr);
        cs[0] = ctx;
        for (int r = 1; r < G; r++) tc_or_die(tc_ctx_create(o.device + r, &cs[r]), "tc_ctx_create (--gpus)");
        vector<string> errs(G);
        vector<std::thread> th;
        for (int r = 1; r < G; r++)
            th.emplace_back([&, r]() {
                vector<const uint64_t *> rows(g.size());
                vector<uint64_t> sizes(g.size());
                for (size_t i = 0; i < g.size(); i++) { rows[i] = g[i].hashes.data(); sizes[i] = g[i].hashes.size(); }
                if (tc_sketches_upload_rows(cs[r], rows.data(), sizes.data(), (uint32_t)g.size())) errs[r] = string("sketch upload, ") + tc_last_error();
            });
        for (std::thread &t : th) t.join();
        for (int r = 0; r < G; r++) if (!errs[r].empty()) die("GPU " + std::to_string(o.device + r) + ": " + errs[r]);
        tc_or_die(tc_greedy_cluster_multi(cs.data(), G, o.kmer, o.containment, o.threshold, o.fast ? 1 : 0, rep_of.data(), &nclust), "greedy clustering");
        for (int r = 1; r < G; r++) tc_ctx_destroy(cs[r]);
    } else
    if (o.fast) tc_or_die(tc_greedy_cluster_kssd(ctx, o.kmer, o.threshold, rep_of.data(), &nclust), "greedy clustering");
    else tc_or_die(tc_greedy_cluster(ctx, o.kmer, o.containment, o.threshold, rep_of.data(), &nclust), "greedy clustering");
    vector<vector<int>> cluster;
    vector<int> cid(g.size(), -1);
    for (int j = 0; j < (int)g.size(); j++) {
        if (rep_of[j] == j) { cid[j] = (int)cluster.size(); cluster.push_back({j}); }
        else cluster[cid[rep_of[j]]].push_back(j);
    }
    if (!o.fast && o.save_rep && (save || o.presketched)) {
        // MinHashInitialClusterWithState + state.save (sub_command.cpp:2673-2679, 2904-2910)
        MinHashState st;
        st.threshold = o.threshold; st.kmer_size = o.kmer; st.sketch_size = (int)g[0].ssize; st.containment = o.containment;
        st.clusters = cluster;
        for (const vector<int> &c : cluster) st.rep_ids.push_back(c[0]);
        const string state_file = folder + "/cluster_state.bin";
        if (save_minhash_state(st, g, state_file)) cerr << "-----saved cluster state (with inverted index) to: " << state_file << endl;
    }
    if (o.fast && o.save_rep && (save || o.presketched)) {
        // KssdInitialClusterWithState + state.save (sub_command.cpp:1962-1968, 2485-2490): the clusters just built, in the
        // sorted sketch order, as the starting point of later clust-greedy --fast --append runs on this folder
        KssdState st;
        st.threshold = o.threshold; st.kmer_size = o.kmer; st.params = g_kssd;
        st.all = g; st.clusters = cluster;
        for (const vector<int> &c : cluster) st.rep_ids.push_back(c[0]);
        const string state_file = folder + "/cluster_state.bin";
        if (save_kssd_state(st, state_file)) cerr << "-----saved cluster state (with inverted index) to: " << state_file << endl;
    }
    print_result(cluster, g, by_file, o.output, -1.0);
    cerr << "-----write the cluster result into: " << o.output << endl;
    cerr << "-----the cluster number of " << o.output << " is: " << cluster.size() << endl;
    cerr << "========time of greedyCluster is: " << now_s() - t2 << "========" << endl;
#else
    if (o.threshold_only) {
        // connected components of the candidate pairs within the threshold, straight from the dense pass
        // (device union-find epilogue; no edge list, no MST, O(n) memory).  Same partition and cluster order (by smallest
        // member id) as the MST route; members are listed in ascending id order instead of BFS order.
        const uint32_t n = (uint32_t)g.size();
        vector<int32_t> labels(n);
        uint32_t nc = 0;
        if (o.gpus > 1) {
            const int G = o.gpus;
            vector<tc_ctx *> cs(G, nullptr);
            cs[0] = ctx;
            for (int r = 1; r < G; r++) tc_or_die(tc_ctx_create(o.device + r, &cs[r]), "tc_ctx_create (--gpus)");
            vector<vector<int32_t>> part(G, vector<int32_t>(n));
            vector<string> errs(G);
            vector<std::thread> th;
            for (int r = 0; r < G; r++)
                th.emplace_back([&, r]() {
                    if (r > 0) {
                        vector<const uint64_t *> rows(n);
                        vector<uint64_t> sizes(n);
                        for (size_t i = 0; i < n; i++) { rows[i] = g[i].hashes.data(); sizes[i] = g[i].hashes.size(); }
                        if (tc_sketches_upload_rows(cs[r], rows.data(), sizes.data(), n)) { errs[r] = string("sketch upload, ") + tc_last_error(); return; }
                    }
                    if (tc_cluster_threshold(cs[r], o.kmer, o.containment, o.threshold, (uint32_t)r, (uint32_t)G, part[r].data(), nullptr))
                        errs[r] = string("threshold clustering, ") + tc_last_error();
                });
            for (std::thread &t : th) t.join();
            for (int r = 0; r < G; r++) if (!errs[r].empty()) die("GPU " + std::to_string(o.device + r) + ": " + errs[r]);
            for (int r = 1; r < G; r++) tc_ctx_destroy(cs[r]);
            labels = part[0];
            for (int r = 1; r < G; r++) tc_or_die(tc_labels_merge(labels.data(), part[r].data(), n, &nc), "label merge");
            if (G == 1) for (uint32_t v = 0; v < n; v++) nc += labels[v] == (int32_t)v;
        } else
            tc_or_die(tc_cluster_threshold(ctx, o.kmer, o.containment, o.threshold, 0, 1, labels.data(), &nc), "threshold clustering");
        cerr << "========time of generateMST is: " << now_s() - t2 << "========" << endl;
        vector<int> cid(n, -1);
        vector<vector<int>> cluster;
        for (uint32_t v = 0; v < n; v++) {
            const int32_t l = labels[v];                // smallest id of the cluster: it is met first
            if (cid[l] < 0) { cid[l] = (int)cluster.size(); cluster.emplace_back(); }
            cluster[cid[l]].push_back((int)v);
        }
        print_result(cluster, g, by_file, o.output, o.threshold);
        cerr << "-----write the cluster result into: " << o.output << endl;
        cerr << "-----the cluster number of: " << o.output << " is: " << cluster.size() << endl;
        (void)save; (void)folder;
        return;
    }
    vector<tc_edge> mst(g.size());
    uint64_t m = 0;
    const double t_mst = now_s();
    if (o.gpus > 1) {
        // Multi-GPU MST: one context and one host thread per device.  Every device holds all sketches
        // (uploaded over its own PCIe link), evaluates its share of the work units of the dense pass, and the Boruvka
        // min-reductions run inside the kernel over NVLink peer memory (tc_mg_*).  Every rank gets the same MST.
        const int G = o.gpus;
        vector<tc_ctx *> cs(G, nullptr);
        cs[0] = ctx;
        for (int r = 1; r < G; r++) tc_or_die(tc_ctx_create(o.device + r, &cs[r]), "tc_ctx_create (--gpus)");
        vector<unsigned char> handles((size_t)G * TC_MG_HANDLE_BYTES);
        for (int r = 0; r < G; r++) tc_or_die(tc_mg_export(cs[r], (uint32_t)g.size(), handles.data() + (size_t)r * TC_MG_HANDLE_BYTES), "tc_mg_export");
        vector<string> errs(G);
        vector<std::thread> th;
        vector<tc_edge> other(g.size());
        std::mutex other_mu;
        for (int r = 0; r < G; r++)
            th.emplace_back([&, r]() {
                auto fail = [&](const char *what) { errs[r] = string(what) + ", " + tc_last_error(); };
                if (tc_mg_join(cs[r], r, G, handles.data())) { fail("tc_mg_join"); return; }
                if (r > 0) {                     // rank 0's sketches are resident already
                    vector<const uint64_t *> rows(g.size());
                    vector<uint64_t> sizes(g.size());
                    for (size_t i = 0; i < g.size(); i++) { rows[i] = g[i].hashes.data(); sizes[i] = g[i].hashes.size(); }
                    if (tc_sketches_upload_rows(cs[r], rows.data(), sizes.data(), (uint32_t)g.size())) { fail("sketch upload"); return; }
                }
                uint64_t mm = 0;
                if (r == 0) { if (tc_mst(cs[r], o.kmer, o.containment, o.threshold, mst.data(), &mm)) { fail("MST generation"); return; } m = mm; }
                else {
                    vector<tc_edge> mine(g.size());
                    if (tc_mst(cs[r], o.kmer, o.containment, o.threshold, mine.data(), &mm)) { fail("MST generation"); return; }
                }
            });
        for (std::thread &t : th) t.join();
        for (int r = 0; r < G; r++) if (!errs[r].empty()) die("GPU " + std::to_string(o.device + r) + ": " + errs[r]);
        for (int r = 0; r < G; r++) tc_mg_leave(cs[r]);
        for (int r = 1; r < G; r++) tc_ctx_destroy(cs[r]);
        if (o.dense) {                           // the density epilogue reads ALL candidate pairs: one full pass on device 0
            uint64_t np = 0, ne = 0;
            tc_or_die(tc_allpairs_pass(ctx, tc_radio(o.threshold, o.kmer), 0, 1, &np, &ne), "dense pass for --dense");
        }
    } else
    tc_or_die(tc_mst(ctx, o.kmer, o.containment, o.threshold, mst.data(), &m), "MST generation");
    mst.resize(m);
    double t3 = now_s();
    if (getenv("TCLUST_TIMING")) cerr << "[tclust timing] tc_mst " << (t3 - t_mst) * 1e3 << " ms" << endl;
    cerr << "========time of generateMST is: " << t3 - t2 << "========" << endl;
    if (save) save_mst(g, mst, folder, by_file, o.fast);
    write_tree_outputs(g, mst, by_file, o);
    vector<vector<int>> cluster = clusters_from_forest(mst, o.threshold, (int)g.size());
    print_result(cluster, g, by_file, o.output, o.threshold);
    cerr << "-----write the cluster result into: " << o.output << endl;
    cerr << "-----the cluster number of: " << o.output << " is: " << cluster.size() << endl;
    if (getenv("TCLUST_TIMING")) cerr << "[tclust timing] forest, clusters, output " << (now_s() - t3) * 1e3 << " ms" << endl;
    if (o.dense) {
        // density maps + ANI histogram over the candidate pairs (GPU epilogue), then the reference's noise-removal
        // pass (sub_command.cpp:3070-3100; getNoiseNode MST.cpp:189-211, modifyForest :87-106)
        double t4 = now_s();
        const int span = 100, n = (int)g.size();                        // DENSE_SPAN, common.hpp:26
        vector<int32_t> dense((size_t)span * n);
        uint64_t ani[101];
        tc_or_die(tc_dense_ani(ctx, o.kmer, o.containment, span, dense.data(), ani), "density / ANI histogram");
        if (save) {
            FILE *fa = fopen((folder + "/mst.ani").c_str(), "w+");
            if (!fa) die("saveANI(), cannot open file: " + folder + "/mst.ani");
            fwrite(ani, sizeof(uint64_t), 101, fa); fclose(fa);
            cerr << "-----save the ani file into: " << folder << "/mst.ani" << endl;
            FILE *fd = fopen((folder + "/mst.dense").c_str(), "w+");
            if (!fd) die("saveDense(), cannot open the file: " + folder + "/mst.dense");
            wr(fd, n); wr(fd, span);
            fwrite(dense.data(), sizeof(int), (size_t)span * n, fd); fclose(fd);
            cerr << "-----save the dense file into: " << folder << endl;
        }
        const int alpha = 2;
        const int dense_index = o.threshold / 0.01;
        vector<char> noise(n, 0);
        size_t total_noise = 0;
        for (const vector<int> &c : cluster) {
            if (c.size() == 1) continue;
            vector<int> vals;
            for (int v : c) vals.push_back(dense[(size_t)dense_index * n + v]);
            std::sort(vals.begin(), vals.end());
            int thr = std::max(std::min(vals[vals.size() / 4] - 1, alpha), 0);
            for (int v : c) if (dense[(size_t)dense_index * n + v] <= thr) { noise[v] = 1; total_noise++; }
        }
        cerr << "-----the total noiseArr size is: " << total_noise << endl;
        vector<tc_edge> forest;
        for (const tc_edge &e : mst) if (e.dist <= o.threshold && !noise[e.pre] && !noise[e.suf]) forest.push_back(e);
        vector<vector<int>> clean = clusters_from_forest(forest, o.threshold, n);
        const string out2 = o.output + ".removeNoise";
        print_result(clean, g, by_file, out2, -1.0);
        cerr << "-----write the cluster without noise into: " << out2 << endl;
        cerr << "-----the cluster number of: " << out2 << " is: " << clean.size() << endl;
        cerr << "========time of tuning cluster is: " << now_s() - t4 << "========" << endl;
    }
#endif
}

void usage(const char *prog)
{
    fprintf(stderr,
            "usage: %s [-t threads] [-m minLen] [-c containCompress | -s sketchSize] [-k kmerSize] [-l] [-e]\n"
            "          [-d threshold] -o output (-i input | --presketched DIR"
#ifndef GREEDY_CLUST
            " | --premsted DIR"
#endif
            ") [--append FILE] [--fast [--drlevel N]] [--dense] [--save-rep]\n"
            "          [--newick-tree] [--phylip-tree] [--nexus-tree] [--linkage-matrix] [--device N] [--gpus G] [--threshold-only]\n", prog);
}

}  // namespace

int main(int argc, char **argv)
{
    Options o;
    if (o.threads < 1) o.threads = 1;
    // The host packer of the 2-bit transport (tc_sketch_* packs on spare cores when there are >= 12 per GPU) would take its
    // cores from the file parsers here, and one GPU's PCIe link (~40 GB/s) is never what limits this feeder (~9 GB/s):
    // measured 3.7 s packed vs 1.7 s plain on 6 Gbp.  The library default still applies to callers with resident buffers.
    setenv("TCLUST_SKETCH_PACK", "0", 0);
    for (int a = 1; a < argc; a++) {
        string f = argv[a];
        auto need = [&](const char *what) -> const char * { if (a + 1 >= argc) die(string("option ") + what + " needs a value"); return argv[++a]; };
        if (f == "-t" || f == "--threads") o.threads = atoi(need("-t"));
        else if (f == "-m" || f == "--min-length") { o.min_len = strtoull(need("-m"), nullptr, 10); fprintf(stderr, "-----set the filter minimum length: %ld\n", (long)o.min_len); }
        else if (f == "-c" || f == "--containment") { o.compress = atoi(need("-c")); o.containment = true;
            fprintf(stderr, "-----use AAF distance with containment coefficient, the sketch size is in porportion with 1/%d\n", o.compress); }
        else if (f == "-k" || f == "--kmer-size") { o.kmer = atoi(need("-k")); o.set_kmer = true; fprintf(stderr, "-----set kmerSize: %d\n", o.kmer); }
        else if (f == "-s" || f == "--sketch-size") { o.sketch_size = atoi(need("-s")); o.jaccard = true; fprintf(stderr, "-----set sketchSize:  %d\n", o.sketch_size); }
        else if (f == "-l" || f == "--list") o.by_file = true;
        else if (f == "-e" || f == "--no-save") o.no_save = true;
        else if (f == "-d" || f == "--threshold") { o.threshold = atof(need("-d")); o.has_threshold = true; fprintf(stderr, "-----set threshold:  %g\n", o.threshold); }
        else if (f == "-o" || f == "--output") { o.output = need("-o"); o.has_output = true; }
        else if (f == "-i" || f == "--input") { o.input = need("-i"); o.has_input = true; }
        else if (f == "--presketched") { o.folder = need("--presketched"); o.presketched = true; }
        else if (f == "--append") { o.input = need("--append"); o.append = true; }
        else if (f == "--save-rep") o.save_rep = true;
#ifndef GREEDY_CLUST
        else if (f == "--premsted") { o.folder = need("--premsted"); o.premsted = true; }
#endif
        else if (f == "--fast") o.fast = true;
        else if (f == "--dense") o.dense = true;
        else if (f == "--newick-tree") o.newick = true;
        else if (f == "--phylip-tree") o.phylip = true;
        else if (f == "--nexus-tree") o.nexus = true;
        else if (f == "--linkage-matrix") o.linkage = true;
        else if (f == "--drlevel") o.drlevel = atoi(need("--drlevel"));
        else if (f == "--device") o.device = atoi(need("--device"));
        else if (f == "--gpus") o.gpus = atoi(need("--gpus"));
        else if (f == "--threshold-only") o.threshold_only = true;
        else if (f == "-h" || f == "--help") { usage(argv[0]); return 0; }
        else { usage(argv[0]); die("unsupported option " + f + " (see DESIGN.md: out of scope for the B200 hot-path build)"); }
    }
    if (!o.has_output) { cerr << "ERROR: option -o/--output is required (unless --buildDB or --stats is used)" << endl; return 1; }
    if (o.threads < 1) { fprintf(stderr, "-----Invalid thread number %d\n", o.threads); return 1; }
    if (o.gpus < 1 || o.gpus > 8) { fprintf(stderr, "-----Invalid --gpus %d (1..8)\n", o.gpus); return 1; }
    fprintf(stderr, "-----set the thread number %d\n", o.threads);
    if (!o.has_threshold) cerr << "-----use default threshold: " << o.threshold << endl;
    if (o.fast && (o.drlevel < 0 || o.drlevel > 8)) { cerr << "ERROR: invalid drlevel " << o.drlevel << ", should be in [0, 8]" << endl; return 1; }
#ifdef GREEDY_CLUST
#else
    if (o.save_rep)
        die("clust-mst --save-rep is not part of this build: the MST medoid state (src/mst_state.cpp) is not built; see DESIGN.md section 8");
#endif

#ifndef GREEDY_CLUST
    if (o.premsted && !o.append) {          // clust_from_mst: host only, no GPU needed
        vector<Genome> g; vector<tc_edge> mst;
        bool by_file = load_mst(g, mst, o.folder, o.fast);
        write_tree_outputs(g, mst, by_file, o);
        vector<vector<int>> cluster = clusters_from_forest(mst, o.threshold, (int)g.size());
        // clust_from_mst prints without the threshold header (sub_command.cpp:1898), clust_from_mst_fast with it (:1790)
        print_result(cluster, g, by_file, o.output, o.fast ? o.threshold : -1.0);
        cerr << "-----write the cluster result into: " << o.output << endl;
        cerr << "-----the cluster number of: " << o.output << " is: " << cluster.size() << endl;
        return 0;
    }
#endif
    // The CUDA context takes about a second to come up on a 180 GB device: it is created on a second thread, so that
    // --presketched can read its sketch files in the meantime; every other flow waits for it right away.
    tc_ctx *ctx = nullptr;
    const bool timing = getenv("TCLUST_TIMING") != nullptr;
    const double t_ctx = now_s();
    string ctx_err;
    std::future<int> ctx_ready = std::async(std::launch::async, [&]() {
        const int rc = tc_ctx_create(o.device, &ctx);
        if (rc) ctx_err = tc_last_error();                  // the error text is per thread
        return rc;
    });
    auto need_ctx = [&]() {
        if (!ctx_ready.valid()) return;
        if (ctx_ready.get()) die("tc_ctx_create: " + ctx_err);
        if (timing) cerr << "[tclust timing] CUDA context ready " << (now_s() - t_ctx) * 1e3 << " ms after start" << endl;
    };
    if (o.append || o.premsted) need_ctx();          // the other flows wait for it where they first need it
    vector<Genome> g;
    if (o.append) {
        // classic append (no saved cluster state): append_clust_mst (sub_command.cpp:1578-1757) and
        // append_clust_greedy (:22-96).  The folder's sketch parameters are reused, tune_parameters is not called.
        if (o.dense) die("--append with --dense is not part of this build");
#ifdef GREEDY_CLUST
        if (o.fast) {
            if (!o.presketched) { cerr << "ERROR option --append, option --presketched needed" << endl; return 1; }
            const int rc = append_greedy_fast(ctx, o);
            tc_ctx_destroy(ctx);
            return rc;
        }
#else
        if (o.fast) {
            // append_clust_mst_fast, the path without a saved mst_cluster_state.bin (sub_command.cpp:1317-1477): KSSD
            // sketches of the new genomes with the folder's half_k / drlevel, new-vs-all edges, Kruskal over the old MST
            // and the new forest.  kssd.hash.sketch keeps the OLD folder's parameter block, genome count included
            // (saveKssdSketches(final_sketches, pre_info, ...), :1349).
            if (!o.presketched && !o.premsted) { cerr << "ERROR: option --append, option --presketched or --premsted needed" << endl; return 1; }
            if (!o.by_file) die("--fast needs -l/--list in this build (SURVEY.md 8a-9)");
            vector<Genome> pre;
            const bool pre_by_file = load_kssd_sketches(pre, o.folder);
            const KssdParams pre_info = g_kssd;
            if (pre_by_file != o.by_file) {
                cerr << "Warning: append_clust_mst(), the input format of append genomes and pre-sketched genome is not same (single input genome vs. genome list)" << endl;
                cerr << "the output cluster file may not have the genome file name" << endl;
            }
            vector<Genome> mst_info; vector<tc_edge> pre_mst;
            load_mst(mst_info, pre_mst, o.folder, true);
            o.kmer = 2 * pre_info.half_k;
            o.drlevel = pre_info.drlevel;
            cerr << "---the thread number is: " << o.threads << endl;
            cerr << "---the threshold is: " << o.threshold << endl;
            vector<Genome> add;
            sketch_files([&]() { return ctx; }, o, add);                                   // sets g_kssd from o.kmer / o.drlevel
            if (g_kssd.half_k != pre_info.half_k || g_kssd.half_subk != pre_info.half_subk || g_kssd.drlevel != pre_info.drlevel)
                die("append_clust_mst_fast(), the KSSD parameters of " + o.folder + " cannot be reproduced for the new genomes");
            g_kssd = pre_info;
            const uint32_t n_pre = (uint32_t)pre.size();
            if (mst_info.size() != n_pre) die("append_clust_mst_fast(), kssd.info.mst and kssd.info.sketch of " + o.folder + " differ in size");
            g = std::move(pre);
            for (Genome &x : add) g.push_back(std::move(x));
            for (size_t i = 0; i < g.size(); i++) g[i].id = (int)i;
            const string folder = current_date_time();
            const bool save = !o.no_save;
            if (save) {
                if (system(("mkdir -p " + folder).c_str()) != 0) die("cannot create " + folder);
                save_kssd_sketches(g, folder, o.by_file);                // with kssd.sketch.index / .dict over all genomes
            }
            upload(ctx, g);
            vector<tc_edge> mst(g.size());
            uint64_t m = 0;
            tc_or_die(tc_mst_append(ctx, o.kmer, o.containment, o.threshold, n_pre, pre_mst.data(), pre_mst.size(), mst.data(), &m),
                      "incremental MST generation");
            mst.resize(m);
            write_tree_outputs(g, mst, pre_by_file, o);
            vector<vector<int>> cluster = clusters_from_forest(mst, o.threshold, (int)g.size());
            print_result(cluster, g, pre_by_file, o.output, o.threshold);
            cerr << "-----write the cluster result into: " << o.output << endl;
            cerr << "-----the cluster number of: " << o.output << " is: " << cluster.size() << endl;
            if (save) save_mst(g, mst, folder, o.by_file, true);
            tc_ctx_destroy(ctx);
            return 0;
        }
#endif
#ifdef GREEDY_CLUST
        if (!o.presketched) { cerr << "ERROR option --append, option --presketched needed" << endl; return 1; }
        if (append_greedy_minhash_state(ctx, o)) { tc_ctx_destroy(ctx); return 0; }      // a saved cluster_state.bin: incremental
#else
        if (!o.presketched && !o.premsted) { cerr << "ERROR option --append, option --presketched or --premsted needed" << endl; return 1; }
#endif
        vector<Genome> pre;
        const bool pre_by_file = load_sketches(pre, o.folder, o);
        cerr << "-----use the same sketch parameters with pre-generated sketches" << endl;
        cerr << "---the kmer size is: " << o.kmer << endl;
        if (o.containment) cerr << "---use the AAF distance (variable-sketch-size), the sketch size is in proportion with 1/" << o.compress << endl;
        else cerr << "---use the Mash distance (fixed-sketch-size), the sketch size is: " << o.sketch_size << endl;
        cerr << "---the thread number is: " << o.threads << endl;
        cerr << "---the threshold is: " << o.threshold << endl;
        vector<Genome> add;
        if (o.by_file) sketch_files([&]() { return ctx; }, o, add); else sketch_sequences(ctx, o, add);
        const uint32_t n_pre = (uint32_t)pre.size();
        g = std::move(pre);
        for (Genome &x : add) g.push_back(std::move(x));
        const string folder = current_date_time();
        const bool save = !o.no_save;
#ifdef GREEDY_CLUST
        if (pre_by_file != o.by_file) {
            cerr << "ERROR: append_clust_greedy(), the input format of append genomes and pre-sketched genome is not same (single input genome vs. genome list)" << endl;
            return 1;
        }
        for (size_t i = 0; i < g.size(); i++) g[i].id = i < n_pre ? g[i].id : (int)(i - n_pre);   // ids as loaded / as sketched
        std::sort(g.begin(), g.end(), [](const Genome &a, const Genome &b) {                    // cmpGenomeSize / cmpSeqSize
            if (a.total_len != b.total_len) return a.total_len > b.total_len;
            return a.id < b.id;
        });
        if (save) {
            if (system(("mkdir -p " + folder).c_str()) != 0) die("cannot create " + folder);
            save_sketches(g, folder, o);
        }
        upload(ctx, g);
        vector<int32_t> centre(g.size());
        uint32_t nclust = 0;
        tc_or_die(tc_greedy_cluster_bruteforce(ctx, o.kmer, (uint32_t)o.sketch_size, o.containment, o.threshold, centre.data(), &nclust),
                  "greedy clustering");
        vector<vector<int>> cluster;
        vector<int> cid(g.size(), -1);
        for (int j = 0; j < (int)g.size(); j++) {
            if (centre[j] == j) { cid[j] = (int)cluster.size(); cluster.push_back({j}); }
            else cluster[cid[centre[j]]].push_back(j);
        }
        print_result(cluster, g, o.by_file, o.output, -1.0);
        cerr << "-----write the cluster result into: " << o.output << endl;
        cerr << "-----the cluster number of " << o.output << " is: " << cluster.size() << endl;
#else
        if (pre_by_file != o.by_file) {
            cerr << "Warning: append_clust_mst(), the input format of append genomes and pre-sketched genome is not same (single input genome vs. genome list)" << endl;
            cerr << "the output cluster file may not have the genome file name" << endl;
        }
        vector<Genome> mst_info; vector<tc_edge> pre_mst;
        load_mst(mst_info, pre_mst, o.folder, false);
        if (mst_info.size() != n_pre) die("append_clust_mst(), info.mst and info.sketch of " + o.folder + " differ in size");
        for (size_t i = 0; i < g.size(); i++) g[i].id = (int)i;
        if (save) {
            if (system(("mkdir -p " + folder).c_str()) != 0) die("cannot create " + folder);
            save_sketches(g, folder, o);
        }
        upload(ctx, g);
        vector<tc_edge> mst(g.size());
        uint64_t m = 0;
        tc_or_die(tc_mst_append(ctx, o.kmer, o.containment, o.threshold, n_pre, pre_mst.data(), pre_mst.size(), mst.data(), &m),
                  "incremental MST generation");
        mst.resize(m);
        write_tree_outputs(g, mst, pre_by_file, o);
        vector<vector<int>> cluster = clusters_from_forest(mst, o.threshold, (int)g.size());
        print_result(cluster, g, pre_by_file, o.output, o.threshold);
        cerr << "-----write the cluster result into: " << o.output << endl;
        cerr << "-----the cluster number of: " << o.output << " is: " << cluster.size() << endl;
        if (save) save_mst(g, mst, folder, o.by_file, false);
#endif
    } else if (o.presketched) {                    // clust_from_sketches (sub_command.cpp:2647-2856)
        double t0 = now_s();
        bool by_file;
        if (o.fast) { by_file = load_kssd_sketches(g, o.folder); o.kmer = 2 * g_kssd.half_k; }   // sub_command.cpp: kmer_size = half_k * 2
        else by_file = load_sketches(g, o.folder, o);
        cerr << "-----the size of sketches is: " << g.size() << endl;
#ifdef GREEDY_CLUST
        if (!o.fast)
        std::sort(g.begin(), g.end(), [](const Genome &a, const Genome &b) {      // cmpGenomeSize / cmpSeqSize
            if (a.total_len != b.total_len) return a.total_len > b.total_len;
            return a.id < b.id;
        });
#endif
        cerr << "========time of load genome Infos and sketch Infos is: " << now_s() - t0 << endl;
        need_ctx();
        run_clustering(ctx, g, o, by_file, false, o.folder);
    } else {
        if (!o.has_input) { usage(argv[0]); die("option -i/--input or --presketched is required"); }
        if (!(o.fast ? tune_kssd_parameters(o) : tune_parameters(o))) return 1;
        double t0 = now_s();
        if (o.by_file) sketch_files([&]() { need_ctx(); return ctx; }, o, g); else { need_ctx(); sketch_sequences(ctx, o, g); }
        need_ctx();
        if (o.fast) o.kmer = 2 * g_kssd.half_k;                          // SketchInfo.cpp:1019-1020, MST.cpp:219
        cerr << "-----the size of sketches (number of genomes or sequences) is: " << g.size() << endl;
        double t1 = now_s();
        cerr << "========time of computing sketch is: " << t1 - t0 << "========" << endl;
        string folder = current_date_time();
        const bool save = !o.no_save;
        if (save) {
            string cmd = "mkdir -p " + folder;
            if (system(cmd.c_str()) != 0) die("cannot create " + folder);
            if (o.fast) save_kssd_sketches(g, folder, o.by_file); else save_sketches(g, folder, o);
            cerr << "========time of saveSketches is: " << now_s() - t1 << "========" << endl;
        }
        run_clustering(ctx, g, o, o.by_file, save, folder);
    }
    const double t_end = now_s();
    need_ctx();
    tc_ctx_destroy(ctx);
    if (timing) cerr << "[tclust timing] context teardown " << (now_s() - t_end) * 1e3 << " ms" << endl;
    return 0;
}
