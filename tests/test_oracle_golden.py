"""Pin the CPU oracle (oracle/tclust_oracle.c) to the reference: golden vectors produced by the
unmodified reference binaries (tools/make_golden.py), plus live comparisons against oracle/_ref
when those binaries are present.  CPU only."""
import glob
import os
import subprocess

import numpy as np
import pytest

from rabbittclust_b200 import fasta, sketch_io, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def genome_records(name):
    for d in ("genomes", "ecoli", "mst_genomes"):
        for p in glob.glob(os.path.join(GOLD, d, name + "*.fa.gz")):
            return [r[2] for r in fasta.read_fasta(p)]
    raise FileNotFoundError(name)


def partition(clusters):
    return sorted(sorted(n for _, n in c) for c in clusters)


def test_murmur_known_answers(oracle):
    # MurmurHash3_x64_128 reference vectors (seed 0 empty = 0; published test values) + Mash-style k-mer
    assert oracle.murmur_h1(b"", 0) == 0
    assert oracle.murmur_h1(b"hello", 0) == 0xcbd8a7b341bd9b02
    assert oracle.murmur_h1(b"The quick brown fox jumps over the lazy dog", 0) == 0xe34bbc7bbc071b6c


@pytest.mark.parametrize("k,s", [(21, 1000), (15, 200), (32, 64), (17, 1000)])
def test_sketch_matches_reference_golden(oracle, k, s):
    gold = np.load(os.path.join(GOLD, "sketches.npz"))
    names = sorted({key.split("_k")[0] for key in gold.files})
    assert len(names) == 8
    for nm in names:
        want = gold["%s_k%d_s%d" % (nm, k, s)]
        got = oracle.minhash_sketch(genome_records(nm), k, s)
        assert np.array_equal(got, want), (nm, k, s)


def test_ecoli_known_answer_from_survey(oracle):
    # SURVEY.md 8c pins the full genomes; the 150 kb slices are pinned by the golden file, and the
    # slice sketches of K-12 strains 1 and 3 share almost everything while O157:H7 does not.
    gold = np.load(os.path.join(GOLD, "sketches.npz"))
    a, b, c = (gold["ecoli%d_k21_s1000" % i] for i in (1, 2, 3))
    c12, _ = oracle.intersect(a, b)
    c13, _ = oracle.intersect(a, c)
    assert c13 > c12
    assert len(a) == 1000 and (np.diff(a.astype(object)) > 0).all()


def test_mst_list_golden(oracle):
    d = os.path.join(GOLD, "mst_list")
    hs = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    by_file, info = sketch_io.read_info_sketch(os.path.join(d, "info.sketch"))
    assert by_file and hs["k"] == 15 and not hs["containment"] and hs["param"] == 1000
    # sketches in the folder == oracle sketches of the genomes named in info.sketch
    for rec, sk in zip(info, hs["sketches"]):
        got = oracle.minhash_sketch(genome_records(rec["name"]), 15, 1000)
        assert np.array_equal(got, sk), rec["name"]
    ref_mst = sketch_io.read_edge_mst(os.path.join(d, "edge.mst"))
    mst = oracle.mst(hs["sketches"], 15, 0.05)
    assert np.array_equal(np.sort(mst["dist"]), np.sort(ref_mst["dist"]))
    n = len(info)
    _, labels, _ = oracle.clusters_from_mst(mst, n, 0.05)
    mine = sorted(sorted(info[i]["name"] for i in np.nonzero(labels == c)[0]) for c in range(labels.max() + 1))
    assert mine == partition(sketch_io.parse_cluster_file(os.path.join(d, "clusters.txt")))


@pytest.mark.parametrize("tag,contain", [("greedy_list_s", False), ("greedy_list_c", True)])
def test_greedy_list_golden(oracle, tag, contain):
    d = os.path.join(GOLD, tag)
    hs = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    _, info = sketch_io.read_info_sketch(os.path.join(d, "info.sketch"))
    assert hs["containment"] == contain
    for rec, sk in zip(info, hs["sketches"]):
        # -c 100: sketch size = max(fileBytes/100, 100) with the gz trailer's uncompressed size
        s = len(sk) if contain else 1000
        got = oracle.minhash_sketch(genome_records(rec["name"]), 15, s)
        assert np.array_equal(got, sk), rec["name"]
    nc, rep = oracle.greedy(hs["sketches"], 15, 0.05, contain)
    ref = sketch_io.parse_cluster_file(os.path.join(d, "clusters.txt"))
    # reference order: clusters in representative-creation order, members in arrival order
    mine = [[info[r]["name"]] + [info[j]["name"] for j in range(len(rep)) if rep[j] == r and j != r]
            for r in sorted(set(rep.tolist()))]
    assert mine == [[n for _, n in c] for c in ref]
    assert nc == len(ref)


@pytest.mark.parametrize("mode,contain", [("mash", False), ("contain", True)])
def test_presketched_golden(oracle, mode, contain):
    d = os.path.join(GOLD, "presketched", mode)
    hs = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    _, info = sketch_io.read_info_sketch(os.path.join(d, "info.sketch"))
    n = len(info)
    mst = oracle.mst(hs["sketches"], hs["k"], 0.05, contain)
    _, labels, order = oracle.clusters_from_mst(mst, n, 0.05)
    mine = sorted(sorted(info[i]["name"] for i in np.nonzero(labels == c)[0]) for c in range(labels.max() + 1))
    assert mine == partition(sketch_io.parse_cluster_file(os.path.join(d, "clust-mst.clusters.txt")))
    # greedy --presketched sorts by (totalSeqLength desc, id asc) first (sub_command.cpp:2657-2660)
    order = sorted(range(n), key=lambda i: (-info[i]["length"], i))
    sk = [hs["sketches"][i] for i in order]
    nc, rep = oracle.greedy(sk, hs["k"], 0.05, contain)
    ref = sketch_io.parse_cluster_file(os.path.join(d, "clust-greedy.clusters.txt"))
    mine = [[info[order[r]]["name"]] + [info[order[j]]["name"] for j in range(n) if rep[j] == r and j != r]
            for r in sorted(set(rep.tolist()))]
    assert mine == [[nm for _, nm in c] for c in ref]


def test_kssd_golden(oracle):
    import struct
    d = os.path.join(GOLD, "kssd")
    data = open(os.path.join(d, "kssd.hash.sketch"), "rb").read()
    pid, half_k, half_subk, drlevel, ngen = struct.unpack_from("<iiiii", data, 0)
    assert (half_subk, drlevel, ngen) == (6, 3, 16)
    pos, sk = 20, []
    for _ in range(ngen):
        (m,) = struct.unpack_from("<Q", data, pos); pos += 8
        sk.append(np.frombuffer(data, "<u4", m, pos).astype(np.uint64)); pos += 4 * m
    assert pos == len(data)
    for i, want in enumerate(sk):
        got = oracle.kssd_sketch(genome_records("g%02d" % i), 2 * half_k, drlevel)
        assert np.array_equal(got, want), i
    ref_mst = sketch_io.read_edge_mst(os.path.join(d, "edge.mst"))
    mst = oracle.mst(sk, 2 * half_k, 0.05)   # KSSD distance = semantic A with k = 2*half_k (MST.cpp:219)
    assert np.array_equal(np.sort(mst["dist"]), np.sort(ref_mst["dist"]))


def test_semantic_B_against_reference_binary(oracle, tmp_path):
    exe = oracle.ref_bin("ref_harness")
    if exe is None:
        pytest.skip("oracle/_ref not built")
    sk = synth.family_sketches(3, 4, s=400, seed=8)
    sketch_io.write_minhash_folder(str(tmp_path), sk, ["x%d" % i for i in range(len(sk))], k=21, sketch_size=400)
    for i, j in [(0, 1), (0, 5), (2, 3), (7, 7)]:
        out = subprocess.check_output([exe, "dist", str(tmp_path / "hash.sketch"), str(i), str(j)]).split()
        jac, dist = float(out[0]), float(out[1])
        mine = oracle.jaccard_B(sk[i], sk[j], 400)
        assert mine == jac
        assert float(oracle.lib().oc_distance_B(mine, 21)) == dist


def test_reference_cli_accepts_our_folder_live(oracle, tmp_path):
    exe = oracle.ref_bin("clust-mst")
    if exe is None:
        pytest.skip("oracle/_ref not built")
    sk = synth.family_sketches(5, 5, s=200, seed=31)
    names = ["n%02d" % i for i in range(len(sk))]
    folder = tmp_path / "sk"
    sketch_io.write_minhash_folder(str(folder), sk, names, k=21, sketch_size=200)
    subprocess.run([exe, "-t", "1", "--presketched", str(folder), "-d", "0.05", "-o", "out.txt"], cwd=tmp_path,
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref = partition(sketch_io.parse_cluster_file(str(tmp_path / "out.txt")))
    mst = oracle.mst(sk, 21, 0.05)
    _, labels, _ = oracle.clusters_from_mst(mst, len(sk), 0.05)
    mine = sorted(sorted(names[i] for i in np.nonzero(labels == c)[0]) for c in range(labels.max() + 1))
    assert mine == ref


SEMB_CASES = ["eq200", "rag200", "rag64", "eq1000"]


def semb_case(name):
    z = np.load(os.path.join(GOLD, "semb", "semb.npz"))
    off = z[name + "_off"].astype(np.int64)
    sk = [z[name + "_hashes"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    s, k, thr6 = (int(x) for x in z[name + "_param"])
    return z, sk, s, k, thr6 / 1e6


@pytest.mark.parametrize("name", SEMB_CASES)
def test_semantic_B_golden(oracle, name):
    """a-11/a-12/a-16: MinHash::jaccard/distance/containDistance, modifyMST and greedyCluster of the unmodified
    reference (tools/make_golden_semb.py) vs the oracle restatement."""
    z, sk, s, k, thr = semb_case(name)
    pairs = z[name + "_pairs"]
    idx = 0
    for i in range(len(sk)):
        for j in range(i):
            c, den = oracle.common_B(sk[i], sk[j], s)
            assert c / den == pairs[idx, 0]
            assert oracle.distance_B(sk[i], sk[j], s, k) == pairs[idx, 1]
            assert oracle.contain_distance(sk[i], sk[j], k) == pairs[idx, 2]
            idx += 1
    for tag, contain in (("", False), ("_contain", True)):
        mst = oracle.mst_bruteforce(sk, k, s, contain)
        assert np.array_equal(np.sort(mst["dist"]), z[name + "_mst_w" + tag])
        _, rep = oracle.greedy_bruteforce(sk, k, s, thr, contain)
        assert np.array_equal(rep, z[name + "_rep" + tag])


def _read_dense(path):
    import struct
    raw = open(path, "rb").read()
    n, span = struct.unpack_from("<ii", raw, 0)
    return np.frombuffer(raw, "<i4", n * span, 8).reshape(span, n)


@pytest.mark.parametrize("tag,contain", [("mash", False), ("contain", True)])
def test_dense_ani_and_noise_removal_golden(oracle, tag, contain):
    """--dense path: mst.ani / mst.dense of the reference's own run and its noise-removed cluster
    files vs the oracle restatement."""
    d = os.path.join(GOLD, "dense", tag)
    hs = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    _, info = sketch_io.read_info_sketch(os.path.join(d, "info.sketch"))
    dense, ani = oracle.dense_ani(hs["sketches"], hs["k"], 0.05, contain)
    assert np.array_equal(ani, np.fromfile(os.path.join(d, "mst.ani"), "<u8"))
    assert np.array_equal(dense, _read_dense(os.path.join(d, "mst.dense")))
    mst = sketch_io.read_edge_mst(os.path.join(d, "edge.mst"))
    _, labels, _ = oracle.remove_noise_clusters(mst, len(info), 0.05, dense)
    mine = sorted(sorted(info[i]["name"] for i in np.nonzero(labels == c)[0]) for c in range(labels.max() + 1))
    assert mine == partition(sketch_io.parse_cluster_file(os.path.join(d, "clusters.txt.removeNoise")))
    # --presketched run (the containment set has 3 noise genomes)
    p = os.path.join(GOLD, "presketched", tag)
    hs = sketch_io.read_hash_sketch(os.path.join(p, "hash.sketch"))
    _, info = sketch_io.read_info_sketch(os.path.join(p, "info.sketch"))
    n = len(info)
    dense, _ = oracle.dense_ani(hs["sketches"], hs["k"], 0.05, hs["containment"])
    mst = oracle.mst(hs["sketches"], hs["k"], 0.05, hs["containment"])
    _, labels, _ = oracle.remove_noise_clusters(mst, n, 0.05, dense)
    mine = sorted(sorted(info[i]["name"] for i in np.nonzero(labels == c)[0]) for c in range(labels.max() + 1))
    assert mine == partition(sketch_io.parse_cluster_file(os.path.join(GOLD, "dense", "presk_%s.removeNoise" % tag)))


def _read_kssd_hash_sketch(path, nsketch=None):
    import struct
    data = open(path, "rb").read()
    hdr = struct.unpack_from("<iiiii", data, 0)            # id, half_k, half_subk, drlevel, genomeNumber
    pos, sk = 20, []
    while pos < len(data) and (nsketch is None or len(sk) < nsketch):
        (m,) = struct.unpack_from("<Q", data, pos); pos += 8
        sk.append(np.frombuffer(data, "<u4", m, pos).astype(np.uint64)); pos += 4 * m
    assert pos == len(data)
    return hdr, sk


def test_append_golden_kssd(oracle):
    """clust-mst --fast --append of the unmodified reference (tools/make_golden_append.py): 9 genomes clustered, 7
    appended.  The new folder holds all 16 sketches behind the OLD parameter block (genomeNumber stays 9,
    sub_command.cpp:1349), and MST(old MST + new-vs-all edges) has the weights of the all-at-once run and of the oracle."""
    d = os.path.join(GOLD, "append", "kssd")
    hdr_b, sk_b = _read_kssd_hash_sketch(os.path.join(d, "base", "kssd.hash.sketch"))
    hdr_r, sk_r = _read_kssd_hash_sketch(os.path.join(d, "result", "kssd.hash.sketch"))
    assert hdr_b[4] == 9 and len(sk_b) == 9
    assert hdr_r == hdr_b and len(sk_r) == 16                      # old parameter block, genome count included
    _, sk_all = _read_kssd_hash_sketch(os.path.join(GOLD, "kssd", "kssd.hash.sketch"))
    assert all(np.array_equal(a, b) for a, b in zip(sk_r, sk_all))
    assert all(np.array_equal(a, b) for a, b in zip(sk_r, sk_b))
    half_k = hdr_b[1]
    res = sketch_io.read_edge_mst(os.path.join(d, "result", "edge.mst"))
    full = sketch_io.read_edge_mst(os.path.join(GOLD, "kssd", "edge.mst"))
    assert np.array_equal(np.sort(res["dist"]), np.sort(full["dist"]))
    mst = oracle.mst(sk_r, 2 * half_k, 0.05)
    assert np.array_equal(np.sort(mst["dist"]), np.sort(res["dist"]))
    # the oracle's append rule itself: forest of the new-vs-all edges + old MST, Kruskal
    old = sketch_io.read_edge_mst(os.path.join(d, "base", "edge.mst"))
    assert np.array_equal(np.sort(oracle.mst(sk_b, 2 * half_k, 0.05)["dist"]), np.sort(old["dist"]))
    assert partition(sketch_io.parse_cluster_file(os.path.join(d, "result", "app.txt"))) == \
        partition(sketch_io.parse_cluster_file(os.path.join(GOLD, "kssd", "clusters.txt")))


def test_append_golden_minhash(oracle):
    """clust-mst --append of the unmodified reference on the same split (Mash mode, k = 15, s = 1000): folder, weights and
    partition of the appended result equal the all-at-once golden run (tests/golden/mst_list) and the oracle's MST."""
    d = os.path.join(GOLD, "append", "minhash")
    res = sketch_io.read_hash_sketch(os.path.join(d, "result", "hash.sketch"))
    full = sketch_io.read_hash_sketch(os.path.join(GOLD, "mst_list", "hash.sketch"))
    assert res["k"] == full["k"] == 15 and res["param"] == full["param"] == 1000 and len(res["sketches"]) == 16
    assert all(np.array_equal(a, b) for a, b in zip(res["sketches"], full["sketches"]))
    assert open(os.path.join(d, "result", "info.sketch"), "rb").read() == open(os.path.join(GOLD, "mst_list", "info.sketch"), "rb").read()
    e_res = sketch_io.read_edge_mst(os.path.join(d, "result", "edge.mst"))
    e_full = sketch_io.read_edge_mst(os.path.join(GOLD, "mst_list", "edge.mst"))
    assert np.array_equal(np.sort(e_res["dist"]), np.sort(e_full["dist"]))
    assert np.array_equal(np.sort(oracle.mst(res["sketches"], 15, 0.05)["dist"]), np.sort(e_res["dist"]))
    assert partition(sketch_io.parse_cluster_file(os.path.join(d, "result", "app.txt"))) == \
        partition(sketch_io.parse_cluster_file(os.path.join(GOLD, "mst_list", "clusters.txt")))


def _clusters_after_increment(clusters, rep_ids, start, rep_of):
    """KssdIncrementalCluster's bookkeeping (greedy.cpp:1851-1870): a new representative opens an EMPTY cluster."""
    clusters = [list(c) for c in clusters]
    rep_ids = list(rep_ids)
    cid = {r: i for i, r in enumerate(rep_ids)}
    for q, r in enumerate(rep_of, start):
        if int(r) == q:
            cid[q] = len(clusters)
            rep_ids.append(q)
            clusters.append([])
        else:
            clusters[cid[int(r)]].append(q)
    return clusters, rep_ids


def test_append_golden_greedy_kssd_state(oracle):
    """clust-greedy --fast --append --save-rep of the unmodified reference (tools/make_golden_append.py): 9 genomes
    clustered, 4 appended without a state (the state is built and saved), 3 more appended from the saved state.  The
    oracle's restatement of KssdIncrementalCluster reproduces both increments from the saved states, and the reader of
    cluster_state.bin sees a consistent file (representatives, clusters, hash -> representative index)."""
    d = os.path.join(GOLD, "append", "greedy_kssd")
    base = sketch_io.read_kssd_hash_sketch(os.path.join(d, "base", "kssd.hash.sketch"))
    _, base_info = sketch_io.read_kssd_info_sketch(os.path.join(d, "base", "kssd.info.sketch"))
    s1 = sketch_io.read_kssd_cluster_state(os.path.join(d, "step1", "cluster_state.bin"))
    s2 = sketch_io.read_kssd_cluster_state(os.path.join(d, "step2", "cluster_state.bin"))
    assert (s1["kmer_size"], s1["half_k"], s1["drlevel"], s1["threshold"]) == (2 * base["half_k"], base["half_k"], base["drlevel"], 0.05)
    # the state keeps the base sketches in the processing order of the initial clustering (size descending) ...
    by_file = {i["file"]: h for i, h in zip(base_info, base["sketches"])}
    assert [len(x["hashes"]) for x in s1["sketches"][:9]] == sorted((len(h) for h in base["sketches"]), reverse=True)
    assert all(np.array_equal(by_file[x["file"]], x["hashes"]) for x in s1["sketches"][:9])
    # ... followed by the appended ones in list order
    for st, step, lo in ((s1, "step1", 9), (s2, "step2", 13)):
        new = sketch_io.read_kssd_hash_sketch(os.path.join(d, step, "new", "kssd.hash.sketch"))
        assert all(np.array_equal(a, x["hashes"]) for a, x in zip(new["sketches"], st["sketches"][lo:]))
        assert len(new["sketches"]) == len(st["sketches"]) - lo
        # index section: hash -> numbers of the representatives holding it
        want = {}
        for r, gid in enumerate(st["rep_ids"]):
            for h in st["sketches"][gid]["hashes"]:
                want.setdefault(int(h), []).append(r)
        assert want == st["index"]
    # increment 1: clusters of the base run (base.txt) + 4 new genomes
    base_clusters = [[gid for gid, _ in c] for c in sketch_io.parse_cluster_file(os.path.join(d, "base", "base.txt"))]
    seeds = [c[0] for c in base_clusters]
    k = s1["kmer_size"]
    nrep, rep_of = oracle.greedy_incremental_kssd([x["hashes"] for x in s1["sketches"]], k, 0.05, 9, seeds)
    cl, reps = _clusters_after_increment(base_clusters, seeds, 9, rep_of)
    assert (cl, reps, nrep) == (s1["clusters"], s1["rep_ids"], len(s1["rep_ids"]))
    # increment 2 from the saved state
    nrep, rep_of = oracle.greedy_incremental_kssd([x["hashes"] for x in s2["sketches"]], k, s1["threshold"], 13, s1["rep_ids"])
    cl, reps = _clusters_after_increment(s1["clusters"], s1["rep_ids"], 13, rep_of)
    assert (cl, reps, nrep) == (s2["clusters"], s2["rep_ids"], len(s2["rep_ids"]))
    # the cluster files list exactly the clusters of the states (a founder is not a member of its own new cluster)
    for st, step in ((s1, "step1"), (s2, "step2")):
        got = [[gid for gid, _ in c] for c in sketch_io.parse_cluster_file(os.path.join(d, step, "app.txt"))]
        assert got == st["clusters"]


def test_greedy_incremental_oracle_properties(oracle):
    """the restated rule on a synthetic set: starting from no representatives it is a plain greedy pass in which every
    sketch is within the threshold of its representative, and splitting the run in two gives the same result."""
    rng = np.random.default_rng(11)
    sk = []
    for f in range(8):
        pool = rng.choice(1 << 24, 400, replace=False)
        for m in range(8):
            keep = rng.choice(pool, int(400 * rng.uniform(0.6, 1.0)), replace=False)
            sk.append(np.unique(np.concatenate([keep, rng.integers(0, 1 << 24, 20)])).astype(np.uint64))
    order = rng.permutation(len(sk))
    sk = [sk[i] for i in order]
    n = len(sk)
    nrep, rep = oracle.greedy_incremental_kssd(sk, 18, 0.05, 0, [])
    reps = [q for q in range(n) if rep[q] == q]
    assert nrep == len(reps) >= 8
    for q in range(n):
        if rep[q] != q:
            assert rep[q] < q and rep[rep[q]] == rep[q]
            c = len(np.intersect1d(sk[q], sk[rep[q]]))
            assert oracle.lib().oc_greedy_dist(c, len(sk[rep[q]]), len(sk[q]), 0, 18) <= 0.05
    half = n // 2
    reps_half = [q for q in reps if q < half]
    nrep2, rep2 = oracle.greedy_incremental_kssd(sk, 18, 0.05, half, reps_half)
    assert nrep2 == nrep and np.array_equal(rep2, rep[half:])


def _multi_fasta_records(path):
    recs, name, cur = [], None, []
    for line in open(path, "rb").read().split(b"\n"):
        if line.startswith(b">"):
            if name is not None:
                recs.append((name, b"".join(cur)))
            name, cur = line[1:].split(b" ")[0].decode(), []
        else:
            cur.append(line)
    recs.append((name, b"".join(cur)))
    return recs


@pytest.mark.parametrize("k", [13, 15, 16])
def test_kssd_single_file_consumer_golden(oracle, k):
    """clust-mst --fast -i of the unmodified reference (tools/make_golden_kssd_seq.py): per-record KSSD sketches of one
    multi-FASTA at --drlevel 0 (every window visible).  k = 13 and 15 are ODD: one stale base in the first window after
    every reset and the un-rounded shift; k = 16 equals the list-mode loop."""
    d = os.path.join(GOLD, "kssd_seq")
    gold = sketch_io.read_kssd_hash_sketch(os.path.join(d, "k%d" % k, "kssd.hash.sketch"))
    by_file, info = sketch_io.read_kssd_info_sketch(os.path.join(d, "k%d" % k, "kssd.info.sketch"))
    assert not by_file and gold["half_k"] == (k + 1) // 2 and gold["drlevel"] == 0
    recs = dict(_multi_fasta_records(os.path.join(d, "multi.fa")))
    assert len(info) == len(recs) == len(gold["sketches"])
    nonempty = 0
    for i, g in zip(info, gold["sketches"]):
        seq = recs[i["name"]]
        assert i["length"] == len(seq)
        mine = oracle.kssd_sketch_seq(seq, k, 0)
        assert np.array_equal(mine, g), (i["name"], len(mine), len(g))
        nonempty += len(g) > 0
        if k % 2 == 0:
            assert np.array_equal(mine, oracle.kssd_sketch([seq], k, 0))
    assert nonempty >= 6
    if k % 2:
        # the quirk is visible: the list-mode rule on the same records gives different sets
        assert any(not np.array_equal(oracle.kssd_sketch([recs[i["name"]]], k, 0), g) for i, g in zip(info, gold["sketches"]))


@pytest.mark.parametrize("tag,contain", [("s", False), ("c", True)])
def test_append_golden_greedy_minhash_state(oracle, tag, contain):
    """clust-greedy --save-rep on 9 genomes, then --append --save-rep of 7 from the saved state (append_clust_greedy,
    sub_command.cpp:97-190; MinHashIncrementalCluster, greedy.cpp:1963-2130), both by the unmodified reference
    (tools/make_golden_append.py).  The oracle's restatement reproduces the increment from the saved state."""
    import json
    d = os.path.join(GOLD, "append", "greedy_minhash", tag)
    s0 = json.load(open(os.path.join(d, "state_base.json")))
    s1 = json.load(open(os.path.join(d, "state_app.json")))
    base = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    assert base["containment"] == contain and s0["containment"] == contain and s0["kmer_size"] == base["k"] == 15
    assert len(base["sketches"]) == 9 == len(s0["sketches"]) and len(s1["sketches"]) == 16
    # the state of the base run: the clusters of base.txt, first member = representative
    base_clusters = [[gid for gid, _ in c] for c in sketch_io.parse_cluster_file(os.path.join(d, "base.txt"))]
    assert s0["clusters"] == base_clusters and s0["rep_ids"] == [c[0] for c in base_clusters]
    # all 16 sketches at these parameters: the committed all-at-once runs of the same genome list
    full = sketch_io.read_hash_sketch(os.path.join(GOLD, "greedy_list_%s" % tag, "hash.sketch"))
    by_len = {}
    _, info = sketch_io.read_info_sketch(os.path.join(GOLD, "greedy_list_%s" % tag, "info.sketch"))
    sk_of = {i["file"]: h for i, h in zip(info, full["sketches"])}
    allsk = [sk_of[x[3]] for x in s1["sketches"]]
    assert [len(h) for h in allsk] == [x[2] for x in s1["sketches"]]
    assert all(np.array_equal(a, b) for a, b in zip(allsk[:9], base["sketches"]))
    nrep, rep_of = oracle.greedy_incremental_minhash(allsk, 15, s0["threshold"], 9, s0["rep_ids"], contain)
    cl, reps = _clusters_after_increment(s0["clusters"], s0["rep_ids"], 9, rep_of)
    assert (cl, reps, nrep) == (s1["clusters"], s1["rep_ids"], len(s1["rep_ids"]))
    got = [[gid for gid, _ in c] for c in sketch_io.parse_cluster_file(os.path.join(d, "app.txt"))]
    assert got == s1["clusters"]


def test_phmap_iteration_order_golden(oracle, tmp_path):
    """the oracle's model of the reference's candidate table (phmap::flat_hash_map<int,int>, reserve(1000)) against
    iteration orders recorded from the real container (tools/make_golden_incremental.py: ref_harness phorder), growth past
    the reserved capacity included; and live against the container when oracle/_ref is built"""
    import json
    for case in json.load(open(os.path.join(GOLD, "incremental", "phmap_order.json"))):
        assert oracle.phmap_iteration_order(case["keys"]).tolist() == case["order"], len(case["keys"])
    exe = oracle.ref_bin("ref_harness")
    if exe is not None:
        rng = np.random.default_rng(77)
        for n in (1, 17, 1000, 1792, 1793, 4000):
            keys = rng.choice(1 << 20, n, replace=False).astype(np.int32)
            keys.tofile(str(tmp_path / "k.bin"))
            subprocess.check_call([exe, "phorder", str(tmp_path / "k.bin"), str(tmp_path / "o.bin")])
            assert oracle.phmap_iteration_order(keys).tolist() == np.fromfile(str(tmp_path / "o.bin"), np.int32).tolist()


def incremental_golden(name):
    """-> (sketches in state order, n_old, seeds, base clusters, golden rep ids, golden clusters, generator args)"""
    import json
    g = json.load(open(os.path.join(GOLD, "incremental", name + ".json")))
    sk, n_old = synth.incremental_tie_set(**g["generator"])
    assert n_old == g["n_old"] and sorted(g["order"]) == list(range(len(sk)))
    sk = [sk[i] for i in g["order"]]
    seeds = [r for r in g["rep_ids"] if r < n_old]
    assert g["rep_ids"][:len(seeds)] == seeds
    base = [[m for m in cl if m < n_old] for cl in g["clusters"][:len(seeds)]]
    return sk, n_old, seeds, base, g["rep_ids"], g["clusters"], g["generator"]


@pytest.mark.parametrize("name", ["kssd", "minhash_s", "minhash_c"])
def test_incremental_tie_rule_golden(oracle, name):
    """KssdIncrementalCluster / MinHashIncrementalCluster of the unmodified reference on sets full of exact distance ties
    (two to four representatives at the same distance; every query touches every representative, so the candidate table
    grows past its reserved capacity): the oracle picks the same representative every time."""
    sk, n_old, seeds, base, want_reps, want_clusters, gen = incremental_golden(name)
    if name == "kssd":
        nrep, rep_of = oracle.greedy_incremental_kssd(sk, gen["kmer"], gen["threshold"], n_old, seeds)
    else:
        nrep, rep_of = oracle.greedy_incremental_minhash(sk, gen["kmer"], gen["threshold"], n_old, seeds, gen.get("containment", False))
    cl, reps = _clusters_after_increment(base, seeds, n_old, rep_of)
    assert reps == want_reps and nrep == len(want_reps)
    assert cl == want_clusters
    # the set does contain ties that the creation-order rule would decide differently
    ties = 0
    for q, r in enumerate(rep_of, n_old):
        if int(r) != q:
            c = len(np.intersect1d(sk[q], sk[int(r)]))
            same = [s for s in reps if s < q and s != int(r) and len(sk[s]) == len(sk[int(r)]) and len(np.intersect1d(sk[q], sk[s])) == c]
            ties += any(s < int(r) for s in same)
    assert ties >= 5, ties


def kssd_greedy_goldens():
    import json
    return json.load(open(os.path.join(GOLD, "incremental", "greedy_kssd_full.json")))


@pytest.mark.parametrize("case", range(3))
def test_greedy_kssd_rule_golden(oracle, case):
    """KssdGreedyClusterWithInvertedIndex through the unmodified clust-greedy --fast --presketched (-t 1) on sets with many
    equal sizes and equal Jaccard values (tools/make_golden_incremental.py): the oracle's restatement gives the same
    clusters in the same order, members in arrival order."""
    g = kssd_greedy_goldens()[case]
    sk = synth.kssd_like_families(**g["generator"])
    sk = [sk[i] for i in g["order"]]
    assert all(len(sk[i]) >= len(sk[i + 1]) for i in range(len(sk) - 1))          # the reference's processing order
    nc, rep = oracle.greedy_kssd(sk, g["kmer"], g["threshold"])
    clusters, cid = [], {}
    for j, r in enumerate(rep):
        if int(r) == j:
            cid[j] = len(clusters)
            clusters.append([j])
        else:
            clusters[cid[int(r)]].append(j)
    assert nc == len(g["clusters"]) and clusters == g["clusters"]


@pytest.mark.parametrize("k,drlevel", [(14, 0), (16, 1), (18, 3), (20, 2)])
def test_kssd_per_record_rule_equals_list_rule_for_even_k(oracle, k, drlevel):
    """for an even k the single-file consumer (tuple not cleared at a gap, window after k valid bases) selects exactly the
    windows of the list-mode loop: random records with gaps of every length, lower case and other codes"""
    rng = np.random.default_rng(k * 10 + drlevel)
    for trial in range(6):
        b = bytearray(synth.random_genome(int(rng.integers(2000, 30000)), rng))
        for p in rng.integers(0, len(b) - 300, 25):
            L = int(rng.choice([1, 2, k - 1, k, k + 1, 60, 250]))
            b[p:p + L] = b"N" * L
        for p in rng.integers(0, len(b), 10):
            b[p] = b"RYKMnx-"[int(rng.integers(0, 7))]
        if trial % 2:
            b[100:900] = bytes(b[100:900]).lower()
        rec = bytes(b)
        assert np.array_equal(oracle.kssd_sketch_seq(rec, k, drlevel), oracle.kssd_sketch([rec], k, drlevel))
    # and for an odd k the two rules differ on a record with gaps
    rec = bytes(synth.random_genome(4000, rng)[:2000]) + b"N" + bytes(synth.random_genome(2000, rng))
    assert not np.array_equal(oracle.kssd_sketch_seq(rec, k - 1, 0), oracle.kssd_sketch([rec], k - 1, 0))
