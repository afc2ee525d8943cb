"""GPU parity of the greedy representative scan (K5) against the CPU oracle (oc_greedy_inverted,
pinned to the reference's clust-greedy -t 1 golden runs) and against those golden runs directly."""
import os

import numpy as np
import pytest

from rabbittclust_b200 import sketch_io, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_scan_common_and_first_hit(ctx, oracle):
    sk = synth.family_sketches(5, 20, s=700, seed=13)
    sk += synth.family_sketches(3, 10, seed=14, ragged=(50, 3000))
    sk.append(np.zeros(0, np.uint64))
    n = len(sk)
    ctx.upload(sk)
    reps = [0, 3, 21, 40, 41, 99, 100, 115, 129, n - 1, 7]          # arbitrary order, not sorted
    common, first = ctx.greedy_scan(10, n, reps)
    for qi, q in enumerate(range(10, n)):
        for ri, r in enumerate(reps):
            c, fh = oracle.intersect(sk[q], sk[r])
            assert common[qi, ri] == c, (q, r)
            assert first[qi, ri] == (fh if c else 0xFFFFFFFF), (q, r)


def _check_greedy(ctx, oracle, sk, k, thr, contain):
    ctx.upload(sk)
    nc, rep = ctx.greedy_cluster(k, thr, contain)
    nc_o, rep_o = oracle.greedy(sk, k, thr, contain)
    assert nc == nc_o
    assert np.array_equal(rep, rep_o)
    return nc


@pytest.mark.parametrize("batch", ["32", "96", "4096"])
def test_greedy_fixed_size_vs_oracle(ctx, oracle, monkeypatch, batch):
    monkeypatch.setenv("TCLUST_GREEDY_BATCH", batch)
    # keep 30-100 % of the pool: some members fall below common_min and found their own clusters
    sk = synth.family_sketches(12, 25, s=1000, seed=5, keep_lo=0.3, keep_hi=1.0)
    rng = np.random.default_rng(1)
    sk = [sk[i] for i in rng.permutation(len(sk))]
    nc = _check_greedy(ctx, oracle, sk, 21, 0.05, False)
    assert 12 <= nc < len(sk)


@pytest.mark.parametrize("contain", [False, True])
@pytest.mark.parametrize("batch", ["64", "4096"])
def test_greedy_ragged_vs_oracle(ctx, oracle, monkeypatch, contain, batch):
    monkeypatch.setenv("TCLUST_GREEDY_BATCH", batch)
    sk = synth.family_sketches(10, 20, seed=8, ragged=(100, 4000), keep_lo=0.3, keep_hi=1.0)
    # greedy --presketched order: size descending (sub_command.cpp:2657-2660)
    sk = sorted(sk, key=lambda s: -len(s))
    _check_greedy(ctx, oracle, sk, 21, 0.05, contain)


def test_greedy_ties_follow_touched_order(ctx, oracle, monkeypatch):
    # many exact ties: identical sketches and representatives that share the same number of hashes
    monkeypatch.setenv("TCLUST_GREEDY_BATCH", "32")
    rng = np.random.default_rng(3)
    pool = np.unique(rng.integers(0, 2**62, 3000, dtype=np.uint64))[:2400]
    a, b, c = pool[:800], pool[800:1600], pool[1600:2400]
    mk = lambda *parts: np.sort(np.concatenate(parts))
    sk = [mk(a[:400], b[:400]), mk(a[400:], c[:400]), mk(b[400:], c[400:])]       # three unrelated reps
    for _ in range(40):
        x, y = rng.choice(3, 2, replace=False)
        srcs = [sk[x], sk[y]]
        sk.append(mk(rng.choice(srcs[0], 400, replace=False), rng.choice(srcs[1], 400, replace=False)))
        sk.append(sk[-1].copy())
    _check_greedy(ctx, oracle, sk, 21, 0.05, False)
    _check_greedy(ctx, oracle, sk, 21, 0.05, True)


@pytest.mark.parametrize("tag,contain", [("greedy_list_s", False), ("greedy_list_c", True)])
def test_greedy_matches_reference_run(ctx, tag, contain):
    d = os.path.join(GOLD, tag)
    hs = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    _, info = sketch_io.read_info_sketch(os.path.join(d, "info.sketch"))
    ctx.upload(hs["sketches"])
    nc, rep = ctx.greedy_cluster(15, 0.05, contain)
    ref = sketch_io.parse_cluster_file(os.path.join(d, "clusters.txt"))
    mine = [[info[r]["name"]] + [info[j]["name"] for j in range(len(rep)) if rep[j] == r and j != r]
            for r in sorted(set(rep.tolist()))]
    assert mine == [[n for _, n in c] for c in ref]
    assert nc == len(ref)


@pytest.mark.parametrize("mode,contain", [("mash", False), ("contain", True)])
def test_greedy_presketched_matches_reference_run(ctx, mode, contain):
    d = os.path.join(GOLD, "presketched", mode)
    hs = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    _, info = sketch_io.read_info_sketch(os.path.join(d, "info.sketch"))
    n = len(info)
    order = sorted(range(n), key=lambda i: (-info[i]["length"], i))
    ctx.upload([hs["sketches"][i] for i in order])
    nc, rep = ctx.greedy_cluster(hs["k"], 0.05, contain)
    ref = sketch_io.parse_cluster_file(os.path.join(d, "clust-greedy.clusters.txt"))
    mine = [[info[order[r]]["name"]] + [info[order[j]]["name"] for j in range(n) if rep[j] == r and j != r]
            for r in sorted(set(rep.tolist()))]
    assert mine == [[nm for _, nm in c] for c in ref]


def test_greedy_large_vs_oracle(ctx, oracle):
    # 3000 genomes, default batch: several batches, representative table grows across groups
    h, off = synth.family_sketches_fast(60, 50, s=1000, seed=17)
    rng = np.random.default_rng(4)
    perm = rng.permutation(3000)
    sk = [h[int(off[i]):int(off[i + 1])] for i in perm]
    nc = _check_greedy(ctx, oracle, sk, 21, 0.05, False)
    assert nc >= 60


def test_greedy_kssd_rule_live_against_reference(ctx, oracle, tmp_path):
    """KssdGreedyClusterWithInvertedIndex of the unmodified reference (clust-greedy --fast --presketched, -t 1) on a
    synthetic KSSD folder with many equal-size ties vs tc_greedy_cluster_kssd on the same processing order."""
    import struct
    import subprocess
    from rabbittclust_b200 import sketch_io
    exe = oracle.ref_bin("clust-greedy")
    if exe is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    sk = []
    for f in range(12):
        pool = rng.choice(1 << 24, 480, replace=False)
        for m in range(10):
            keep = rng.choice(pool, int(480 * rng.uniform(0.5, 1.0)), replace=False)
            priv = rng.integers(0, 1 << 24, int(rng.integers(0, 60)))
            sk.append(np.unique(np.concatenate([keep, priv])).astype(np.uint64))
    n = len(sk)
    d = tmp_path / "kssd"
    d.mkdir()
    half_k, half_subk, drlevel = 9, 6, 3
    with open(d / "kssd.hash.sketch", "wb") as f:
        f.write(struct.pack("<iiiii", (half_k << 8) + (half_subk << 4) + drlevel, half_k, half_subk, drlevel, n))
        for h in sk:
            f.write(struct.pack("<Q", len(h))); f.write(h.astype("<u4").tobytes())
    with open(d / "kssd.info.sketch", "wb") as f:
        f.write(struct.pack("<?Q", True, n))
        for i in range(n):
            fn, nm, cm = b"f%03d.fna" % i, b"s%03d" % i, b"c"
            f.write(struct.pack("<iiiiQ", len(fn), len(nm), len(cm), 0, 2000000)); f.write(fn + nm + cm); f.write(struct.pack("<?", False))
    subprocess.run([exe, "-t", "1", "--fast", "--presketched", str(d), "-d", "0.05", "-o", "out.txt"], cwd=tmp_path, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    # the cluster file carries (position in the reference's processing order, name): recover that order (its
    # std::sort by size is not stable) and replay exactly the same sequence on the GPU path
    ref = sketch_io.parse_cluster_file(str(tmp_path / "out.txt"))
    pos_of = {}
    for c in ref:
        for gid, nm in c:
            pos_of[gid] = int(nm[1:])
    assert sorted(pos_of) == list(range(n))
    order = [pos_of[p] for p in range(n)]
    assert all(len(sk[order[i]]) >= len(sk[order[i + 1]]) for i in range(n - 1))
    ctx.upload([sk[i] for i in order])
    nc, rep = ctx.greedy_cluster_kssd(18, 0.05)
    want = np.zeros(n, np.int32)
    for c in ref:
        for gid, _ in c:
            want[gid] = c[0][0]
    assert nc == len(ref)
    assert np.array_equal(rep, want)


@pytest.mark.parametrize("mode", ["fixed", "ragged-contain", "kssd"])
def test_greedy_sharded_representatives(tc, mode, monkeypatch):
    """tc_greedy_cluster_multi: the representative table dealt over several contexts (here on one device, on a
    multi-GPU box one per device) gives the clusters of the single-context run, whatever the number of contexts."""
    import torch
    monkeypatch.setenv("TCLUST_GREEDY_BATCH", "64")        # many batches: the tables grow on every context
    if mode == "fixed":
        sk = synth.family_sketches(12, 25, s=400, seed=41)
        kw = dict(containment=False, kssd=False)
    elif mode == "ragged-contain":
        sk = synth.family_sketches(10, 20, seed=42, ragged=(100, 1500))
        sk.sort(key=lambda a: -len(a))
        kw = dict(containment=True, kssd=False)
    else:
        sk = synth.family_sketches(10, 20, seed=43, ragged=(200, 600))
        sk.sort(key=lambda a: -len(a))
        kw = dict(containment=False, kssd=True)
    with tc.Context(0) as c:
        c.upload(sk)
        want = c.greedy_cluster_kssd(21, 0.05) if mode == "kssd" else c.greedy_cluster(21, 0.05, kw["containment"])
    ndev = torch.cuda.device_count()
    for g in (2, 3):
        ctxs = [tc.Context(r % ndev) for r in range(g)]
        try:
            for c in ctxs:
                c.upload(sk)
            got = tc.greedy_cluster_multi(ctxs, 21, 0.05, **kw)
        finally:
            for c in ctxs:
                c.close()
        assert got[0] == want[0] and np.array_equal(got[1], want[1]), (mode, g)


def _kssd_like_families(seed, n_fam=12, per_fam=10):
    rng = np.random.default_rng(seed)
    sk = []
    for f in range(n_fam):
        pool = rng.choice(1 << 24, 480, replace=False)
        for m in range(per_fam):
            keep = rng.choice(pool, int(480 * rng.uniform(0.3, 1.0)), replace=False)      # some pairs of a family fall below the threshold
            priv = rng.integers(0, 1 << 24, int(rng.integers(0, 60)))
            sk.append(np.unique(np.concatenate([keep, priv])).astype(np.uint64))
    return [sk[i] for i in rng.permutation(len(sk))]


@pytest.mark.parametrize("batch", ["32", "4096"])
@pytest.mark.parametrize("split", [0, 1, 50, 119])
def test_greedy_incremental_kssd_vs_oracle(ctx, oracle, monkeypatch, batch, split):
    """tc_greedy_incremental_kssd (KssdIncrementalCluster, greedy.cpp:1736-1880) against the oracle's restatement (pinned
    to the reference's saved states, tests/test_oracle_golden.py): the first `split` sketches are clustered by the oracle
    from scratch, their representatives seed the increment over the rest -- new representatives inside a batch included."""
    monkeypatch.setenv("TCLUST_GREEDY_BATCH", batch)
    sk = _kssd_like_families(21)
    n = len(sk)
    _, rep0 = oracle.greedy_incremental_kssd(sk[:split], 18, 0.05, 0, [])
    seeds = [q for q in range(split) if rep0[q] == q]
    want_n, want = oracle.greedy_incremental_kssd(sk, 18, 0.05, split, seeds)
    ctx.upload(sk)
    got_n, got = ctx.greedy_incremental_kssd(18, 0.05, split, seeds)
    assert got_n == want_n
    assert np.array_equal(got, want)
    assert got_n >= len(seeds) and (split > 1 or got_n > 12)       # families split: several representatives per family


def test_greedy_incremental_kssd_rejects_bad_seeds(ctx, tc):
    sk = _kssd_like_families(3, 2, 4)
    ctx.upload(sk)
    with pytest.raises(tc.TclustError):
        ctx.greedy_incremental_kssd(18, 0.05, 4, [5])          # a representative must lie before start_index
    with pytest.raises(tc.TclustError):
        ctx.greedy_incremental_kssd(18, 0.05, 4, [1, 1])       # listed twice


@pytest.mark.parametrize("contain", [False, True])
@pytest.mark.parametrize("batch", ["32", "4096"])
def test_greedy_incremental_minhash_vs_oracle(ctx, oracle, monkeypatch, contain, batch):
    """tc_greedy_incremental (MinHashIncrementalCluster, greedy.cpp:1963-2130) against the oracle's restatement (pinned to
    the reference's saved states): fixed-size Mash sketches and ragged containment sketches, seeded with the
    representatives of a plain greedy pass over the first half."""
    monkeypatch.setenv("TCLUST_GREEDY_BATCH", batch)
    if contain:
        sk = synth.family_sketches(10, 20, seed=8, ragged=(100, 4000), keep_lo=0.3, keep_hi=1.0)
    else:
        sk = synth.family_sketches(12, 25, s=1000, seed=5, keep_lo=0.3, keep_hi=1.0)
    rng = np.random.default_rng(2)
    sk = [sk[i] for i in rng.permutation(len(sk))]
    n, half = len(sk), len(sk) // 2
    _, rep0 = oracle.greedy(sk[:half], 21, 0.05, contain)
    seeds = [q for q in range(half) if rep0[q] == q]
    want_n, want = oracle.greedy_incremental_minhash(sk, 21, 0.05, half, seeds, contain)
    ctx.upload(sk)
    got_n, got = ctx.greedy_incremental(21, 0.05, half, seeds, contain)
    assert got_n == want_n
    assert np.array_equal(got, want)
    assert (got != np.arange(half, n)).any() and got_n >= len(seeds)


@pytest.mark.parametrize("batch", ["64", "4096"])
@pytest.mark.parametrize("name", ["kssd", "minhash_s", "minhash_c"])
def test_greedy_incremental_tie_rule_golden(ctx, oracle, monkeypatch, name, batch):
    """the incremental rules on the reference's own runs over sets full of exact distance ties
    (tests/golden/incremental, tools/make_golden_incremental.py): among equal distances the first candidate in the
    iteration order of the reference's phmap candidate table wins -- same representative for every sketch."""
    from test_oracle_golden import _clusters_after_increment, incremental_golden
    monkeypatch.setenv("TCLUST_GREEDY_BATCH", batch)
    sk, n_old, seeds, base, want_reps, want_clusters, gen = incremental_golden(name)
    ctx.upload(sk)
    if name == "kssd":
        nrep, rep_of = ctx.greedy_incremental_kssd(gen["kmer"], gen["threshold"], n_old, seeds)
    else:
        nrep, rep_of = ctx.greedy_incremental(gen["kmer"], gen["threshold"], n_old, seeds, gen.get("containment", False))
    cl, reps = _clusters_after_increment(base, seeds, n_old, rep_of)
    assert reps == want_reps and nrep == len(want_reps)
    assert cl == want_clusters


@pytest.mark.parametrize("case", range(3))
def test_greedy_kssd_rule_golden_and_oracle(ctx, oracle, case):
    """tc_greedy_cluster_kssd against the committed runs of the unmodified clust-greedy --fast --presketched
    (tests/golden/incremental/greedy_kssd_full.json) and against the oracle's restatement of the rule"""
    from test_oracle_golden import kssd_greedy_goldens
    g = kssd_greedy_goldens()[case]
    sk = synth.kssd_like_families(**g["generator"])
    sk = [sk[i] for i in g["order"]]
    ctx.upload(sk)
    nc, rep = ctx.greedy_cluster_kssd(g["kmer"], g["threshold"])
    nc_o, rep_o = oracle.greedy_kssd(sk, g["kmer"], g["threshold"])
    assert nc == nc_o == len(g["clusters"])
    assert np.array_equal(rep, rep_o)
    want = np.zeros(len(sk), np.int32)
    for c in g["clusters"]:
        want[c] = c[0]
    assert np.array_equal(rep, want)
