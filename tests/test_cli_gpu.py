"""End-to-end drop-in check: our C++ host binaries (rabbittclust_b200/host/bin/clust-mst, clust-greedy;
CUDA hot paths through the C ABI) against the UNMODIFIED reference binaries (oracle/_ref, run with -t 1)
on the same synthetic genome files: sketch folder byte-identical, MST weights bit-identical, same
partition / same greedy cluster file."""
import glob
import gzip
import os
import subprocess

import numpy as np
import pytest

from rabbittclust_b200 import sketch_io, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "rabbittclust_b200", "host", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def _need_bins():
    for p in (os.path.join(OURS, "clust-mst"), os.path.join(OURS, "clust-greedy")):
        if not os.path.exists(p):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "rabbittclust_b200", "host")])
    if not os.path.exists(os.path.join(REF, "clust-mst")):
        pytest.skip("oracle/_ref not built")


def _write_fasta(path, records, width=80):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + b"\n")


def _make_genomes(d, n_fam=5, per_fam=6, length=120000, seed=12345):
    gen = synth.genome_families(n_fam, per_fam, length, seed=seed)
    rng = np.random.default_rng(seed)
    files = []
    for i, g in enumerate(gen):
        fam, mem = divmod(i, per_fam)
        # some genomes in two records, some gz, one with lower case and N runs, one without a comment
        cut = int(rng.integers(length // 3, 2 * length // 3))
        if i % 3 == 0:
            recs = [("g%04d_%03d synthetic family %d member %d" % (fam, mem, fam, mem), g[:cut]), ("ctg2 second contig", g[cut:])]
        elif i % 3 == 1:
            recs = [("g%04d_%03d" % (fam, mem), g)]
        else:
            b = bytearray(g.lower() if i % 2 else g)
            b[1000:1040] = b"N" * 40
            recs = [("g%04d_%03d\tfamily %d" % (fam, mem, fam), bytes(b))]
        p = os.path.join(d, "g%04d_%03d.fna%s" % (fam, mem, ".gz" if i % 4 == 1 else ""))
        _write_fasta(p, recs)
        files.append(p)
    lst = os.path.join(d, "list.txt")
    open(lst, "w").write("\n".join(files) + "\n")
    multi = os.path.join(d, "all.fna")
    _write_fasta(multi, [("seq%d len=%d" % (i, len(g)), g) for i, g in enumerate(gen)])
    return lst, multi


def _run(exe, args, cwd):
    os.makedirs(cwd, exist_ok=True)
    r = subprocess.run([exe] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    folders = [p for p in glob.glob(os.path.join(cwd, "20*")) if os.path.isdir(p)]
    return folders[0] if folders else None, r.stderr.decode()


def _partition(path):
    return sorted(sorted(n for _, n in c) for c in sketch_io.parse_cluster_file(path))


def _same_bytes(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


@pytest.fixture(scope="module")
def genomes(tmp_path_factory):
    _need_bins()
    d = str(tmp_path_factory.mktemp("genomes"))
    return _make_genomes(d)


@pytest.mark.parametrize("mode", ["list", "single"])
@pytest.mark.parametrize("extra", [["-s", "1000"], ["-c", "100"]])
def test_clust_mst_matches_reference(genomes, tmp_path, mode, extra):
    lst, multi = genomes
    inp = ["-l", "-i", lst] if mode == "list" else ["-i", multi]
    args = inp + ["-k", "21", "-d", "0.05", "-o", "out.txt"] + extra
    rf, rerr = _run(os.path.join(REF, "clust-mst"), ["-t", "1"] + args, str(tmp_path / "ref"))
    of, oerr = _run(os.path.join(OURS, "clust-mst"), args, str(tmp_path / "ours"))
    assert "replace the kmerSize" in rerr and "replace the kmerSize" in oerr          # k override
    assert _same_bytes(os.path.join(rf, "hash.sketch"), os.path.join(of, "hash.sketch"))
    assert _same_bytes(os.path.join(rf, "info.sketch"), os.path.join(of, "info.sketch"))
    assert _same_bytes(os.path.join(rf, "info.mst"), os.path.join(of, "info.mst"))
    a = sketch_io.read_edge_mst(os.path.join(rf, "edge.mst"))
    b = sketch_io.read_edge_mst(os.path.join(of, "edge.mst"))
    assert np.array_equal(np.sort(a["dist"]), np.sort(b["dist"]))
    ref_out, our_out = str(tmp_path / "ref" / "out.txt"), str(tmp_path / "ours" / "out.txt")
    assert open(ref_out).read().splitlines()[:3] == open(our_out).read().splitlines()[:3]   # header lines
    assert _partition(ref_out) == _partition(our_out)
    # cluster order (by smallest member id) is tree-independent
    first = lambda p: [c[0][0] for c in sketch_io.parse_cluster_file(p)]
    assert first(ref_out) == first(our_out)
    # --presketched and --premsted on the folder the reference wrote
    _run(os.path.join(OURS, "clust-mst"), ["--presketched", rf, "-d", "0.05", "-o", "pre.txt"], str(tmp_path / "ours"))
    assert _partition(str(tmp_path / "ours" / "pre.txt")) == _partition(ref_out)
    # --premsted on the reference's tree, with every tree/linkage output: same tree -> same bytes as the reference's
    # own --premsted run (which prints no threshold header, sub_command.cpp:1898)
    tree_flags = ["--newick-tree", "--phylip-tree", "--nexus-tree", "--linkage-matrix"]
    _run(os.path.join(REF, "clust-mst"), ["-t", "1", "--premsted", rf, "-d", "0.05", "-o", "mst.txt"] + tree_flags, str(tmp_path / "ref"))
    _run(os.path.join(OURS, "clust-mst"), ["--premsted", rf, "-d", "0.05", "-o", "mst.txt"] + tree_flags, str(tmp_path / "ours"))
    for suffix in ("", ".newick.tree", ".phylip.tree", ".nexus.tree", ".linkage.txt"):
        assert open(str(tmp_path / "ours" / ("mst.txt" + suffix))).read() == open(str(tmp_path / "ref" / ("mst.txt" + suffix))).read(), suffix
    assert _partition(str(tmp_path / "ours" / "mst.txt")) == _partition(ref_out)


@pytest.mark.parametrize("mode", ["list", "single"])
@pytest.mark.parametrize("extra", [["-s", "1000"], ["-c", "100"], []])
def test_clust_greedy_matches_reference(genomes, tmp_path, mode, extra):
    lst, multi = genomes
    inp = ["-l", "-i", lst] if mode == "list" else ["-i", multi]
    args = inp + ["-k", "21", "-d", "0.05", "-o", "out.txt"] + extra
    rf, _ = _run(os.path.join(REF, "clust-greedy"), ["-t", "1"] + args, str(tmp_path / "ref"))
    of, _ = _run(os.path.join(OURS, "clust-greedy"), args, str(tmp_path / "ours"))
    assert _same_bytes(os.path.join(rf, "hash.sketch"), os.path.join(of, "hash.sketch"))
    assert _same_bytes(os.path.join(rf, "info.sketch"), os.path.join(of, "info.sketch"))
    ref_out, our_out = str(tmp_path / "ref" / "out.txt"), str(tmp_path / "ours" / "out.txt")
    assert open(ref_out).read() == open(our_out).read()          # -t 1 greedy output is fully deterministic
    # --presketched: both sort by (length desc, id asc) first
    _run(os.path.join(REF, "clust-greedy"), ["-t", "1", "--presketched", rf, "-d", "0.05", "-o", "pre.txt"], str(tmp_path / "ref"))
    _run(os.path.join(OURS, "clust-greedy"), ["--presketched", rf, "-d", "0.05", "-o", "pre.txt"], str(tmp_path / "ours"))
    assert open(str(tmp_path / "ref" / "pre.txt")).read() == open(str(tmp_path / "ours" / "pre.txt")).read()


def test_clust_mst_fast_matches_reference(genomes, tmp_path):
    """--fast (KSSD) list mode: sketch folder byte-identical to the reference's, same MST weights and partition;
    the reference reads the folder we wrote (incl. kssd.sketch.index/.dict) and we read the reference's."""
    lst, _ = genomes
    args = ["--fast", "-l", "-i", lst, "-d", "0.05", "-o", "out.txt"]
    rf, _ = _run(os.path.join(REF, "clust-mst"), ["-t", "1"] + args, str(tmp_path / "ref"))
    of, _ = _run(os.path.join(OURS, "clust-mst"), args, str(tmp_path / "ours"))
    for fn in ("kssd.hash.sketch", "kssd.info.sketch", "kssd.info.mst"):
        assert _same_bytes(os.path.join(rf, fn), os.path.join(of, fn)), fn
    a = sketch_io.read_edge_mst(os.path.join(rf, "edge.mst"))
    b = sketch_io.read_edge_mst(os.path.join(of, "edge.mst"))
    assert np.array_equal(np.sort(a["dist"]), np.sort(b["dist"]))
    ref_out, our_out = str(tmp_path / "ref" / "out.txt"), str(tmp_path / "ours" / "out.txt")
    assert open(ref_out).read().splitlines()[:3] == open(our_out).read().splitlines()[:3]
    assert _partition(ref_out) == _partition(our_out)
    # cross-reading of the folders
    _run(os.path.join(REF, "clust-mst"), ["-t", "1", "--fast", "--presketched", of, "-d", "0.05", "-o", "x.txt"], str(tmp_path / "ref"))
    assert _partition(str(tmp_path / "ref" / "x.txt")) == _partition(ref_out)
    _run(os.path.join(OURS, "clust-mst"), ["--fast", "--presketched", rf, "-d", "0.05", "-o", "pre.txt"], str(tmp_path / "ours"))
    assert _partition(str(tmp_path / "ours" / "pre.txt")) == _partition(ref_out)
    _run(os.path.join(OURS, "clust-mst"), ["--fast", "--premsted", rf, "-d", "0.05", "-o", "mst.txt"], str(tmp_path / "ours"))
    assert open(str(tmp_path / "ours" / "mst.txt")).read() == open(ref_out).read()
    # explicit k and a different reduction level
    args2 = ["--fast", "-l", "-i", lst, "-k", "20", "--drlevel", "2", "-d", "0.05", "-o", "out2.txt"]
    rf2, _ = _run(os.path.join(REF, "clust-mst"), ["-t", "1"] + args2, str(tmp_path / "ref2"))
    of2, _ = _run(os.path.join(OURS, "clust-mst"), args2, str(tmp_path / "ours2"))
    assert _same_bytes(os.path.join(rf2, "kssd.hash.sketch"), os.path.join(of2, "kssd.hash.sketch"))
    assert _partition(str(tmp_path / "ref2" / "out2.txt")) == _partition(str(tmp_path / "ours2" / "out2.txt"))


def test_clust_mst_fast_golden_folder(tmp_path):
    """the committed reference run (tests/golden/kssd): our binary reproduces its kssd.hash.sketch from the same
    genome files without needing the reference binary"""
    _need_bins()
    os.symlink(os.path.join(ROOT, "tests"), str(tmp_path / "tests"))
    of, _ = _run(os.path.join(OURS, "clust-mst"), ["--fast", "-l", "-i", "tests/golden/mst_genomes/list.txt", "-d", "0.05",
                                                    "-o", "clusters.txt"], str(tmp_path))
    gold = os.path.join(ROOT, "tests", "golden", "kssd")
    assert _same_bytes(os.path.join(gold, "kssd.hash.sketch"), os.path.join(of, "kssd.hash.sketch"))
    assert _same_bytes(os.path.join(gold, "kssd.info.sketch"), os.path.join(of, "kssd.info.sketch"))
    a = sketch_io.read_edge_mst(os.path.join(gold, "edge.mst"))
    b = sketch_io.read_edge_mst(os.path.join(of, "edge.mst"))
    assert np.array_equal(np.sort(a["dist"]), np.sort(b["dist"]))
    assert _partition(os.path.join(gold, "clusters.txt")) == _partition(str(tmp_path / "clusters.txt"))


def test_clust_greedy_fast_matches_reference(genomes, tmp_path):
    """clust-greedy --fast (KssdGreedyClusterWithInvertedIndex): same sketch folder, byte-identical cluster
    file with -t 1 (the size sort is the same std::sort on the same key sequence), also from --presketched"""
    lst, _ = genomes
    args = ["--fast", "-l", "-i", lst, "-d", "0.05", "-o", "out.txt"]
    rf, _ = _run(os.path.join(REF, "clust-greedy"), ["-t", "1"] + args, str(tmp_path / "ref"))
    of, _ = _run(os.path.join(OURS, "clust-greedy"), args, str(tmp_path / "ours"))
    for fn in ("kssd.hash.sketch", "kssd.info.sketch"):
        assert _same_bytes(os.path.join(rf, fn), os.path.join(of, fn)), fn
    assert open(str(tmp_path / "ref" / "out.txt")).read() == open(str(tmp_path / "ours" / "out.txt")).read()
    _run(os.path.join(REF, "clust-greedy"), ["-t", "1", "--fast", "--presketched", rf, "-d", "0.02", "-o", "pre.txt"], str(tmp_path / "ref"))
    _run(os.path.join(OURS, "clust-greedy"), ["--fast", "--presketched", rf, "-d", "0.02", "-o", "pre.txt"], str(tmp_path / "ours"))
    assert open(str(tmp_path / "ref" / "pre.txt")).read() == open(str(tmp_path / "ours" / "pre.txt")).read()


def _split_list(lst, tmp_path, n_first):
    files = open(lst).read().split()
    a, b = str(tmp_path / "first.txt"), str(tmp_path / "second.txt")
    open(a, "w").write("\n".join(files[:n_first]) + "\n")
    open(b, "w").write("\n".join(files[n_first:]) + "\n")
    return a, b


@pytest.mark.parametrize("extra", [["-s", "1000"], ["-c", "100"]])
def test_clust_mst_append_matches_reference(genomes, tmp_path, extra):
    """classic --append: the reference clusters the first 20 genomes, then both binaries append the
    other 10 to that folder: new-vs-all pass + Kruskal over the old MST and the new forest."""
    lst, _ = genomes
    first, second = _split_list(lst, tmp_path, 20)
    base, _ = _run(os.path.join(REF, "clust-mst"), ["-t", "1", "-l", "-i", first, "-k", "21", "-d", "0.05", "-o", "base.txt"] + extra,
                   str(tmp_path / "base"))
    # upstream bug: the reference's append path reads minhash.sketch.index (written by its own normal run) with an
    # incompatible loader and aborts with std::bad_array_new_length; without the cache file it rebuilds the index
    idx = os.path.join(base, "minhash.sketch.index")
    if os.path.exists(idx):
        os.remove(idx)
    args = ["-l", "--append", second, "--presketched", base, "-d", "0.05", "-o", "app.txt", "--linkage-matrix"]
    rf, _ = _run(os.path.join(REF, "clust-mst"), ["-t", "1"] + args, str(tmp_path / "ref"))
    of, _ = _run(os.path.join(OURS, "clust-mst"), args, str(tmp_path / "ours"))
    for fn in ("hash.sketch", "info.sketch", "info.mst"):
        assert _same_bytes(os.path.join(rf, fn), os.path.join(of, fn)), fn
    a = sketch_io.read_edge_mst(os.path.join(rf, "edge.mst"))
    b = sketch_io.read_edge_mst(os.path.join(of, "edge.mst"))
    assert np.array_equal(np.sort(a["dist"]), np.sort(b["dist"]))
    ref_out, our_out = str(tmp_path / "ref" / "app.txt"), str(tmp_path / "ours" / "app.txt")
    assert open(ref_out).read().splitlines()[:3] == open(our_out).read().splitlines()[:3]
    assert _partition(ref_out) == _partition(our_out)
    # the appended result equals clustering everything at once
    full, _ = _run(os.path.join(REF, "clust-mst"), ["-t", "1", "-l", "-i", lst, "-k", "21", "-d", "0.05", "-o", "full.txt"] + extra,
                   str(tmp_path / "full"))
    c = sketch_io.read_edge_mst(os.path.join(full, "edge.mst"))
    assert np.array_equal(np.sort(b["dist"]), np.sort(c["dist"]))
    # linkage matrices agree in their distance column (tie order inside equal weights is tree dependent)
    la = [l.split("\t")[2] for l in open(ref_out + ".linkage.txt")]
    lb = [l.split("\t")[2] for l in open(our_out + ".linkage.txt")]
    assert la == lb


@pytest.mark.parametrize("extra", [["-s", "1000"], ["-c", "100"]])
def test_clust_greedy_append_matches_reference(genomes, tmp_path, extra):
    """classic clust-greedy --append: concatenation sorted by genome length, brute-force greedyCluster with
    MinHash::distance / containDistance (semantic B) -- byte-identical cluster file and sketch folder"""
    lst, _ = genomes
    first, second = _split_list(lst, tmp_path, 18)
    base, _ = _run(os.path.join(REF, "clust-greedy"), ["-t", "1", "-l", "-i", first, "-k", "21", "-d", "0.05", "-o", "base.txt"] + extra,
                   str(tmp_path / "base"))
    args = ["-l", "--append", second, "--presketched", base, "-d", "0.05", "-o", "app.txt"]
    rf, _ = _run(os.path.join(REF, "clust-greedy"), ["-t", "1"] + args, str(tmp_path / "ref"))
    of, _ = _run(os.path.join(OURS, "clust-greedy"), args, str(tmp_path / "ours"))
    for fn in ("hash.sketch", "info.sketch"):
        assert _same_bytes(os.path.join(rf, fn), os.path.join(of, fn)), fn
    assert open(str(tmp_path / "ref" / "app.txt")).read() == open(str(tmp_path / "ours" / "app.txt")).read()


def test_clust_mst_fast_append_matches_reference(genomes, tmp_path):
    """clust-mst --fast --append (append_clust_mst_fast without a saved state, sub_command.cpp:1317-1477): the reference
    clusters the first 20 genomes with KSSD sketches, then both binaries append the other 10 to that folder.  The new
    folder's kssd.hash.sketch keeps the old folder's parameter block; index/dict differ only in entry order."""
    lst, _ = genomes
    first, second = _split_list(lst, tmp_path, 20)
    base, _ = _run(os.path.join(REF, "clust-mst"), ["-t", "1", "--fast", "-l", "-i", first, "-d", "0.05", "-o", "base.txt"],
                   str(tmp_path / "base"))
    args = ["--fast", "-l", "--append", second, "--presketched", base, "-d", "0.05", "-o", "app.txt", "--linkage-matrix"]
    rf, _ = _run(os.path.join(REF, "clust-mst"), ["-t", "1"] + args, str(tmp_path / "ref"))
    of, _ = _run(os.path.join(OURS, "clust-mst"), args, str(tmp_path / "ours"))
    for fn in ("kssd.hash.sketch", "kssd.info.sketch", "kssd.info.mst"):
        assert _same_bytes(os.path.join(rf, fn), os.path.join(of, fn)), fn
    a = sketch_io.read_edge_mst(os.path.join(rf, "edge.mst"))
    b = sketch_io.read_edge_mst(os.path.join(of, "edge.mst"))
    assert np.array_equal(np.sort(a["dist"]), np.sort(b["dist"]))
    ref_out, our_out = str(tmp_path / "ref" / "app.txt"), str(tmp_path / "ours" / "app.txt")
    assert open(ref_out).read().splitlines()[:3] == open(our_out).read().splitlines()[:3]
    assert _partition(ref_out) == _partition(our_out)
    # the appended result equals clustering everything at once, and the reference reads the folder we wrote
    full, _ = _run(os.path.join(REF, "clust-mst"), ["-t", "1", "--fast", "-l", "-i", lst, "-d", "0.05", "-o", "full.txt"],
                   str(tmp_path / "full"))
    c = sketch_io.read_edge_mst(os.path.join(full, "edge.mst"))
    assert np.array_equal(np.sort(b["dist"]), np.sort(c["dist"]))
    _run(os.path.join(REF, "clust-mst"), ["-t", "1", "--fast", "--presketched", of, "-d", "0.05", "-o", "x.txt"], str(tmp_path / "ref"))
    assert _partition(str(tmp_path / "ref" / "x.txt")) == _partition(ref_out)
    la = [l.split("\t")[2] for l in open(ref_out + ".linkage.txt")]
    lb = [l.split("\t")[2] for l in open(our_out + ".linkage.txt")]
    assert la == lb


def test_clust_mst_gpus_flag(genomes, tmp_path):
    """clust-mst --gpus G (multi-GPU group inside the binary: one context + host thread per device, Boruvka reductions
    over peer memory): same edge.mst and cluster file as the single-GPU run, also with --dense."""
    import torch
    g = min(torch.cuda.device_count(), 4)
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    lst, _ = genomes
    args = ["-l", "-i", lst, "-k", "21", "-d", "0.05", "-s", "1000", "--dense", "-o", "out.txt"]
    f1, _ = _run(os.path.join(OURS, "clust-mst"), args, str(tmp_path / "one"))
    fg, _ = _run(os.path.join(OURS, "clust-mst"), args + ["--gpus", str(g)], str(tmp_path / "many"))
    for name in ("edge.mst", "hash.sketch", "mst.ani", "mst.dense"):
        assert _same_bytes(os.path.join(f1, name), os.path.join(fg, name)), name
    assert open(str(tmp_path / "one" / "out.txt")).read() == open(str(tmp_path / "many" / "out.txt")).read()


def test_clust_mst_threshold_only(genomes, tmp_path):
    """clust-mst --threshold-only: same clusters (and cluster order, by smallest member) as the MST route, no edge.mst."""
    lst, _ = genomes
    args = ["-l", "-i", lst, "-k", "21", "-d", "0.05", "-s", "1000"]
    f1, _ = _run(os.path.join(OURS, "clust-mst"), args + ["-o", "mst.txt"], str(tmp_path / "a"))
    f2, _ = _run(os.path.join(OURS, "clust-mst"), args + ["-o", "thr.txt", "--threshold-only"], str(tmp_path / "b"))
    a, b = str(tmp_path / "a" / "mst.txt"), str(tmp_path / "b" / "thr.txt")
    assert _partition(a) == _partition(b)
    first = lambda p: [min(i for i, _ in c) for c in sketch_io.parse_cluster_file(p)]
    assert first(a) == first(b)
    assert open(a).read().splitlines()[:3] == open(b).read().splitlines()[:3]
    assert os.path.exists(os.path.join(f1, "edge.mst")) and not os.path.exists(os.path.join(f2, "edge.mst"))


def _read_kssd_hash_sketch(path):
    """kssd.hash.sketch: KssdParameters{id, half_k, half_subk, drlevel, genomeNumber} then per genome size_t n + n hashes
    (u32 unless half_k - drlevel > 8)."""
    import struct
    data = open(path, "rb").read()
    pid, half_k, half_subk, drlevel, ng = struct.unpack_from("<5i", data, 0)
    w = 8 if half_k - drlevel > 8 else 4
    pos, out = 20, []
    while pos < len(data):
        (n,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        out.append(np.frombuffer(data, "<u8" if w == 8 else "<u4", n, pos).astype(np.uint64))
        pos += n * w
    return (pid, half_k, half_subk, drlevel, ng), out


@pytest.mark.parametrize("kflag", [[], ["-k", "17"], ["-k", "15", "--drlevel", "2"]])
def test_clust_mst_fast_single_file_matches_reference(genomes, tmp_path, kflag):
    """--fast with ONE multi-FASTA (-i without -l; reference: sketchSequencesWithKssd / consumer_fasta_task_with_kssd,
    src/SketchInfo.cpp:274-433,554-642): one KSSD sketch per record.  Same parameter block, same hash SET per record (the
    reference stores them in hash-set iteration order, we ascending), kssd.info.* byte-identical, same MST weights and
    partition; each binary reads the folder the other wrote.  The tuned k is even (16); -k 17 and -k 15 take the odd-k
    rule of that consumer (one stale base per reset, un-rounded shift)."""
    _, multi = genomes
    args = ["--fast", "-i", multi, "-d", "0.05", "-o", "out.txt"] + kflag
    rf, rerr = _run(os.path.join(REF, "clust-mst"), ["-t", "1"] + args, str(tmp_path / "ref"))
    of, oerr = _run(os.path.join(OURS, "clust-mst"), args, str(tmp_path / "ours"))
    pr, hr = _read_kssd_hash_sketch(os.path.join(rf, "kssd.hash.sketch"))
    po, ho = _read_kssd_hash_sketch(os.path.join(of, "kssd.hash.sketch"))
    assert pr == po and len(hr) == len(ho) == pr[4]
    for a, b in zip(hr, ho):
        assert np.array_equal(np.sort(a), b), "hash set of a record differs"
    for fn in ("kssd.info.sketch", "kssd.info.mst"):
        assert _same_bytes(os.path.join(rf, fn), os.path.join(of, fn)), fn
    a = sketch_io.read_edge_mst(os.path.join(rf, "edge.mst"))
    b = sketch_io.read_edge_mst(os.path.join(of, "edge.mst"))
    assert np.array_equal(np.sort(a["dist"]), np.sort(b["dist"]))
    ref_out, our_out = str(tmp_path / "ref" / "out.txt"), str(tmp_path / "ours" / "out.txt")
    assert _partition(ref_out) == _partition(our_out)
    _run(os.path.join(REF, "clust-mst"), ["-t", "1", "--fast", "--presketched", of, "-d", "0.05", "-o", "x.txt"], str(tmp_path / "ref"))
    assert _partition(str(tmp_path / "ref" / "x.txt")) == _partition(ref_out)
    _run(os.path.join(OURS, "clust-mst"), ["--fast", "--presketched", rf, "-d", "0.05", "-o", "pre.txt"], str(tmp_path / "ours"))
    assert _partition(str(tmp_path / "ours" / "pre.txt")) == _partition(ref_out)
