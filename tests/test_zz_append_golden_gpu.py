"""Our clust-mst binary against the committed reference runs of the classic --append flows
(tests/golden/append, tools/make_golden_append.py): no reference binary needed at test time."""
import os

import numpy as np
import pytest

from rabbittclust_b200 import sketch_io
from test_cli_gpu import OURS, ROOT, _need_bins, _partition, _run, _same_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["kssd", "minhash"])
def test_clust_mst_append_golden_folders(tmp_path, name):
    """the committed reference runs of tools/make_golden_append.py (9 genomes clustered, 7 appended): our binary appends
    the same genomes to the committed base folder and reproduces the reference's result folder without needing the
    reference binary"""
    _need_bins()
    os.symlink(os.path.join(ROOT, "tests"), str(tmp_path / "tests"))
    gold = os.path.join(ROOT, "tests", "golden", "append", name)
    fast = ["--fast"] if name == "kssd" else []
    of, _ = _run(os.path.join(OURS, "clust-mst"), fast + ["-l", "--append", "tests/golden/append/second.txt", "--presketched",
                                                           "tests/golden/append/%s/base" % name, "-d", "0.05", "-o", "app.txt"], str(tmp_path))
    pre = "kssd." if name == "kssd" else ""
    for fn in (pre + "hash.sketch", pre + "info.sketch", pre + "info.mst"):
        assert _same_bytes(os.path.join(gold, "result", fn), os.path.join(of, fn)), fn
    a = sketch_io.read_edge_mst(os.path.join(gold, "result", "edge.mst"))
    b = sketch_io.read_edge_mst(os.path.join(of, "edge.mst"))
    assert np.array_equal(np.sort(a["dist"]), np.sort(b["dist"]))
    ref_out, our_out = os.path.join(gold, "result", "app.txt"), str(tmp_path / "app.txt")
    assert open(ref_out).read().splitlines()[:3] == open(our_out).read().splitlines()[:3]
    assert _partition(ref_out) == _partition(our_out)


def _state_without_index_order(path):
    s = sketch_io.read_kssd_cluster_state(path)
    s["sketches"] = [(x["id"], x["length"], x["use64"], x["file"], x["hashes"].tolist()) for x in s["sketches"]]
    return s


def test_clust_greedy_fast_append_state_golden(tmp_path):
    """clust-greedy --fast --append --save-rep (append_clust_greedy_fast, sub_command.cpp:192-276) on the committed
    reference runs: increment 1 builds the state from the pre-sketched folder, increment 2 continues from the saved
    cluster_state.bin.  Cluster files byte-identical; state files equal field by field (the hash -> representative index is
    compared as a map: the reference writes it in hash-map iteration order)."""
    _need_bins()
    import shutil
    os.symlink(os.path.join(ROOT, "tests"), str(tmp_path / "tests"))
    gold = os.path.join(ROOT, "tests", "golden", "append", "greedy_kssd")
    base = str(tmp_path / "base")
    shutil.copytree(os.path.join(gold, "base"), base)
    exe = os.path.join(OURS, "clust-greedy")
    for step, lst in (("step1", "second_a.txt"), ("step2", "second_b.txt")):
        _, err = _run(exe, ["--fast", "-l", "--append", "tests/golden/append/" + lst, "--presketched", base, "--save-rep",
                            "-d", "0.05", "-o", step + ".txt"], str(tmp_path))
        assert ("Initial State Building Mode" if step == "step1" else "Incremental Update Mode") in err
        assert _same_bytes(os.path.join(gold, step, "app.txt"), str(tmp_path / (step + ".txt"))), step
        assert _state_without_index_order(os.path.join(gold, step, "cluster_state.bin")) == \
            _state_without_index_order(os.path.join(base, "cluster_state.bin")), step
    # the sketches of the appended genomes were saved to folders of their own, as compute_kssd_sketches does
    new = sorted(p for p in os.listdir(str(tmp_path)) if p.startswith("20"))
    assert len(new) == 2
    for folder, step in zip(new, ("step1", "step2")):
        for fn in ("kssd.hash.sketch", "kssd.info.sketch"):
            assert _same_bytes(os.path.join(gold, step, "new", fn), str(tmp_path / folder / fn)), (step, fn)


def test_clust_greedy_fast_append_without_save_rep(tmp_path):
    """without --save-rep the same result is printed and no state file is left behind (sub_command.cpp:235-243)"""
    _need_bins()
    import shutil
    os.symlink(os.path.join(ROOT, "tests"), str(tmp_path / "tests"))
    gold = os.path.join(ROOT, "tests", "golden", "append", "greedy_kssd")
    base = str(tmp_path / "base")
    shutil.copytree(os.path.join(gold, "base"), base)
    _run(os.path.join(OURS, "clust-greedy"), ["--fast", "-l", "--append", "tests/golden/append/second_a.txt", "--presketched", base,
                                               "-d", "0.05", "-o", "app.txt"], str(tmp_path))
    assert _same_bytes(os.path.join(gold, "step1", "app.txt"), str(tmp_path / "app.txt"))
    assert not os.path.exists(os.path.join(base, "cluster_state.bin"))


def test_reference_continues_from_our_cluster_state_live(tmp_path):
    """live: a larger synthetic set (several candidates per query).  clust-greedy --fast --save-rep writes the state next
    to the sketches; the unmodified reference appends to OUR state file and we append to ITS state file: all four
    combinations give the same cluster file."""
    _need_bins()
    import shutil
    from test_cli_gpu import REF, _make_genomes
    d = str(tmp_path / "genomes")
    os.makedirs(d)
    lst, _ = _make_genomes(d, n_fam=5, per_fam=6, length=600000, seed=77)
    files = open(lst).read().split()
    rng = np.random.default_rng(5)
    files = [files[i] for i in rng.permutation(len(files))]
    first, second = str(tmp_path / "first.txt"), str(tmp_path / "second.txt")
    open(first, "w").write("\n".join(files[:18]) + "\n")
    open(second, "w").write("\n".join(files[18:]) + "\n")
    out = {}
    for who, exe, t1 in (("ref", os.path.join(REF, "clust-greedy"), ["-t", "1"]), ("ours", os.path.join(OURS, "clust-greedy"), [])):
        cwd = str(tmp_path / who)
        folder, _ = _run(exe, t1 + ["--fast", "-l", "-i", first, "-d", "0.05", "--save-rep", "-o", "base.txt"], cwd)
        assert os.path.exists(os.path.join(folder, "cluster_state.bin")), who
        out[who] = folder
    assert _same_bytes(str(tmp_path / "ref" / "base.txt"), str(tmp_path / "ours" / "base.txt"))
    a = _state_without_index_order(os.path.join(out["ref"], "cluster_state.bin"))
    b = _state_without_index_order(os.path.join(out["ours"], "cluster_state.bin"))
    assert a == b
    results = {}
    for state_of in ("ref", "ours"):
        for who, exe, t1 in (("ref", os.path.join(REF, "clust-greedy"), ["-t", "1"]), ("ours", os.path.join(OURS, "clust-greedy"), [])):
            cwd = str(tmp_path / ("app_%s_on_%s" % (who, state_of)))
            folder = os.path.join(cwd, "state")
            os.makedirs(cwd)
            shutil.copytree(out[state_of], folder)
            _run(exe, t1 + ["--fast", "-l", "--append", second, "--presketched", folder, "--save-rep", "-d", "0.05", "-o", "app.txt"], cwd)
            results[(who, state_of)] = open(os.path.join(cwd, "app.txt"), "rb").read()
            assert os.path.getsize(os.path.join(folder, "cluster_state.bin")) > os.path.getsize(os.path.join(out[state_of], "cluster_state.bin"))
    assert len(set(results.values())) == 1, {k: len(v) for k, v in results.items()}


@pytest.mark.parametrize("tag,flags", [("s", ["-s", "1000"]), ("c", ["-c", "100"])])
def test_clust_greedy_minhash_state_golden(tmp_path, tag, flags):
    """clust-greedy --save-rep from genomes, then --append --save-rep from the saved MinHash state (append_clust_greedy,
    sub_command.cpp:97-190) against the committed reference runs: cluster files byte-identical, states equal field by
    field (hashes of the sketches included, through the folder's hash.sketch)."""
    _need_bins()
    import json
    os.symlink(os.path.join(ROOT, "tests"), str(tmp_path / "tests"))
    gold = os.path.join(ROOT, "tests", "golden", "append", "greedy_minhash", tag)
    exe = os.path.join(OURS, "clust-greedy")

    def summary(path):
        st = sketch_io.read_minhash_cluster_state(path)
        want = {}
        for r, gid in enumerate(st["rep_ids"]):
            for h in st["sketches"][gid]["hashes"]:
                want.setdefault(int(h), []).append(r)
        assert want == st["index"]
        return {"threshold": st["threshold"], "kmer_size": st["kmer_size"], "sketch_size": st["sketch_size"],
                "containment": st["containment"], "rep_ids": st["rep_ids"], "clusters": st["clusters"],
                "sketches": [[x["id"], x["length"], len(x["hashes"]), x["file"]] for x in st["sketches"]]}
    folder, _ = _run(exe, ["-l", "-i", "tests/golden/append/first.txt", "-k", "15"] + flags + ["-d", "0.05", "--save-rep", "-o", "base.txt"],
                     str(tmp_path))
    for fn in ("hash.sketch", "info.sketch"):
        assert _same_bytes(os.path.join(gold, fn), os.path.join(folder, fn)), fn
    assert _same_bytes(os.path.join(gold, "base.txt"), str(tmp_path / "base.txt"))
    assert summary(os.path.join(folder, "cluster_state.bin")) == json.load(open(os.path.join(gold, "state_base.json")))
    _, err = _run(exe, ["-l", "--append", "tests/golden/append/second.txt", "--presketched", folder, "--save-rep", "-d", "0.05", "-o", "app.txt"],
                  str(tmp_path))
    assert "Incremental Update Mode (MinHash)" in err
    assert _same_bytes(os.path.join(gold, "app.txt"), str(tmp_path / "app.txt"))
    assert summary(os.path.join(folder, "cluster_state.bin")) == json.load(open(os.path.join(gold, "state_app.json")))


def test_reference_continues_from_our_minhash_state_live(tmp_path):
    """live, a larger set: the reference appends to OUR MinHash state and we append to ITS state -- same cluster file"""
    _need_bins()
    import shutil
    from test_cli_gpu import REF, _make_genomes
    d = str(tmp_path / "genomes")
    os.makedirs(d)
    lst, _ = _make_genomes(d, n_fam=5, per_fam=6, length=150000, seed=78)
    files = open(lst).read().split()
    rng = np.random.default_rng(6)
    files = [files[i] for i in rng.permutation(len(files))]
    first, second = str(tmp_path / "first.txt"), str(tmp_path / "second.txt")
    open(first, "w").write("\n".join(files[:18]) + "\n")
    open(second, "w").write("\n".join(files[18:]) + "\n")
    out = {}
    bins = (("ref", os.path.join(REF, "clust-greedy"), ["-t", "1"]), ("ours", os.path.join(OURS, "clust-greedy"), []))
    for who, exe, t1 in bins:
        folder, _ = _run(exe, t1 + ["-l", "-i", first, "-k", "21", "-s", "1000", "-d", "0.05", "--save-rep", "-o", "base.txt"], str(tmp_path / who))
        assert os.path.exists(os.path.join(folder, "cluster_state.bin")), who
        out[who] = folder
    assert _same_bytes(str(tmp_path / "ref" / "base.txt"), str(tmp_path / "ours" / "base.txt"))
    results = {}
    for state_of in ("ref", "ours"):
        for who, exe, t1 in bins:
            cwd = str(tmp_path / ("app_%s_on_%s" % (who, state_of)))
            folder = os.path.join(cwd, "state")
            os.makedirs(cwd)
            shutil.copytree(out[state_of], folder)
            idx = os.path.join(folder, "minhash.sketch.index")          # cache file of the reference's normal run (upstream bug)
            if os.path.exists(idx):
                os.remove(idx)
            _run(exe, t1 + ["-l", "--append", second, "--presketched", folder, "--save-rep", "-d", "0.05", "-o", "app.txt"], cwd)
            results[(who, state_of)] = open(os.path.join(cwd, "app.txt"), "rb").read()
    assert len(set(results.values())) == 1, {k: len(v) for k, v in results.items()}


def test_clust_greedy_fast_state_single_file_mode_live(tmp_path, oracle):
    """live, sketch-by-sequence: clust-greedy --fast -i (one multi-FASTA) --save-rep, then --append of a second multi-FASTA
    from the saved state.  The initial clustering equals the unmodified reference's byte for byte, and the rows of
    state-loaded sketches print as the reference prints them (length 0, empty names: its seqInfo is value-initialised).
    The reference's incremental step itself is broken in this mode: it feeds the UNSORTED hash arrays of the single-file
    consumer (hash-set iteration order) to a sorted-merge Jaccard (calculate_mash_distance, greedy.cpp:72-119), so its
    distances are garbage and nearly every new record founds a cluster.  Ours is checked against the oracle instead."""
    _need_bins()
    from test_cli_gpu import REF, _make_genomes, _write_fasta
    from test_oracle_golden import _clusters_after_increment
    from rabbittclust_b200 import fasta
    d = str(tmp_path / "genomes")
    os.makedirs(d)
    _, multi = _make_genomes(d, n_fam=4, per_fam=6, length=300000, seed=91)
    recs = fasta.read_fasta(multi)
    rng = np.random.default_rng(8)
    order = rng.permutation(len(recs))
    first, second = str(tmp_path / "first.fna"), str(tmp_path / "second.fna")
    _write_fasta(first, [("%s %s" % (recs[i][0], recs[i][1] or ""), recs[i][2]) for i in order[:14]])
    _write_fasta(second, [("%s %s" % (recs[i][0], recs[i][1] or ""), recs[i][2]) for i in order[14:]])
    out, folders = {}, {}
    for who, exe, t1 in (("ref", os.path.join(REF, "clust-greedy"), ["-t", "1"]), ("ours", os.path.join(OURS, "clust-greedy"), [])):
        cwd = str(tmp_path / who)
        folder, _ = _run(exe, t1 + ["--fast", "-i", first, "-d", "0.05", "--save-rep", "-o", "base.txt"], cwd)
        assert os.path.exists(os.path.join(folder, "cluster_state.bin")), who
        if who == "ours":
            before = sketch_io.read_kssd_cluster_state(os.path.join(folder, "cluster_state.bin"))
        _run(exe, t1 + ["--fast", "--append", second, "--presketched", folder, "--save-rep", "-d", "0.05", "-o", "app.txt"], cwd)
        out[who] = (open(os.path.join(cwd, "base.txt"), "rb").read(), open(os.path.join(cwd, "app.txt"), "rb").read())
        folders[who] = folder
    assert out["ref"][0] == out["ours"][0]
    # rows of the 14 state-loaded sketches: same text in both files
    def old_rows(text):
        return sorted(l.split(b"\t", 2)[2] for l in text.split(b"\n") if l.startswith(b"\t") and int(l.split(b"\t")[2]) < 14)
    assert old_rows(out["ref"][1]) == old_rows(out["ours"][1]) and len(old_rows(out["ours"][1])) == 14
    # our increment = the oracle's on the sketches of our state
    after = sketch_io.read_kssd_cluster_state(os.path.join(folders["ours"], "cluster_state.bin"))
    sk = [x["hashes"] for x in after["sketches"]]
    nrep, rep_of = oracle.greedy_incremental_kssd(sk, after["kmer_size"], after["threshold"], 14, before["rep_ids"])
    cl, reps = _clusters_after_increment(before["clusters"], before["rep_ids"], 14, rep_of)
    assert (cl, reps) == (after["clusters"], after["rep_ids"])
    got = [[gid for gid, _ in c] for c in sketch_io.parse_cluster_file(str(tmp_path / "ours" / "app.txt"))]
    assert got == after["clusters"]
    assert sum(len(c) for c in cl) > sum(len(c) for c in before["clusters"]) + 3       # new records do join clusters
