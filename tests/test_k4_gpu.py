"""Tile-candidacy flags (k3t_tileflags.cu), the device-resident Boruvka (k4_boruvka.cu) and the multi-GPU group
(tc_mg_*) against the CPU oracle (oracle/tclust_oracle.c: candidate graph + Kruskal, src/MST.cpp:59-75, 1467-1533)."""
import threading

import numpy as np
import pytest

from rabbittclust_b200 import synth

pytestmark = pytest.mark.gpu


def _same_forest(oracle, a, b, n, thr):
    assert len(a) == len(b)
    assert np.array_equal(np.sort(a["dist"]).view(np.uint64), np.sort(b["dist"]).view(np.uint64)), "weights differ (bit-exact)"
    _, la, _ = oracle.clusters_from_mst(a, n, thr)
    _, lb, _ = oracle.clusters_from_mst(b, n, thr)
    part = lambda lab: sorted(sorted(np.nonzero(lab == c)[0].tolist()) for c in range(lab.max() + 1))
    assert part(la) == part(lb)


def _graphy_sketches(seed, n=260, universe=6000, lo=1, hi=90):
    """Small random sketches over a small universe: a sparse candidate graph with many equal weights, chains,
    isolated vertices and components of every size -- everything Boruvka's hooking has to get right."""
    rng = np.random.default_rng(seed)
    sk = [np.unique(rng.integers(0, universe, int(rng.integers(lo, hi)), dtype=np.uint64)) for _ in range(n)]
    sk[3] = np.zeros(0, np.uint64)
    sk[11] = sk[10].copy()
    sk[12] = sk[10].copy()
    return sk


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("containment", [False, True])
def test_device_boruvka_random_graphs(ctx, oracle, seed, containment):
    sk = _graphy_sketches(seed)
    ctx.upload(sk)
    k, thr = 21, 0.3
    got = ctx.mst(k, thr, containment)
    want = oracle.mst(sk, k, thr, containment)
    _same_forest(oracle, got, want, len(sk), thr)
    assert np.all(got["suf"] < got["pre"])
    order = np.lexsort((got["suf"], got["pre"], got["dist"]))
    assert np.array_equal(order, np.arange(len(got))), "forest not sorted by (dist, i, j)"
    st = ctx.mst_stats()
    assert st["rounds"] >= 2 and st["boruvka_ms"] > 0


def test_tile_flags_do_not_change_the_edges(tc, oracle, monkeypatch):
    """Family sets in generation order and shuffled (reordered copy), flags on and off: the same candidate edges."""
    h, off = synth.family_sketches_fast(60, 40, s=400, seed=8)
    n = len(off) - 1
    perm = np.random.default_rng(2).permutation(n)
    hs = np.ascontiguousarray(h.reshape(n, 400)[perm].reshape(-1))
    radio = tc.radio(0.05, 21)
    res = {}
    for name, hh in (("family", h), ("shuffled", hs)):
        for flags in ("1", "0"):
            monkeypatch.setenv("TCLUST_TILE_FLAGS", flags)
            with tc.Context(0) as c:
                c.upload(hashes=hh, off=off)
                c.allpairs_pass(radio)
                e = c.edges()
                st_pass = (len(e), c.mst(21, 0.05))
                stats = c.mst_stats()
            key = np.sort(e["i"].astype(np.uint64) << np.uint64(40) | e["j"].astype(np.uint64) << np.uint64(16) | e["common"].astype(np.uint64))
            res[(name, flags)] = (key, st_pass[1], stats)
        a, b = res[(name, "1")], res[(name, "0")]
        assert np.array_equal(a[0], b[0]), name
        assert np.array_equal(a[1], b[1]), name
        assert a[2]["tiles_flagged"] > 0 and b[2]["tiles_flagged"] < 0
        assert a[2]["tiles_flagged"] < a[2]["tiles_total"] // 4, "family sets must leave most tiles unvisited"
    # the edges of the family order against the oracle's candidate graph
    sk = [h[int(off[i]):int(off[i + 1])] for i in range(n)]
    want = oracle.common_matrix(sk[:200])
    got = tc.Context(0)
    try:
        got.upload(sk[:200])
        assert np.array_equal(got.pair_common(0, 200, 0, 200, impl=3), want)
    finally:
        got.close()


def test_flags_with_unlisted_blocks_and_partitions(tc, monkeypatch):
    """Blocks too large for a union list (no list: their row and column are always visited) and row blocks with several
    table partitions, next to family blocks: flags on == flags off == the K3H-only pass."""
    rng = np.random.default_rng(4)
    h, off = synth.family_sketches_fast(40, 32, s=300, seed=3)
    sk = [h[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    for i in range(64, 96):          # one block of very large sketches: 32 x 12000 > the sort capacity
        sk[i] = np.unique(np.concatenate([sk[i], rng.integers(0, 2**64 - 4, 12000, dtype=np.uint64)]))
    for i in range(320, 352):        # one block with ~30000 distinct shared-with-someone ranks: several partitions
        sk[i] = np.unique(np.concatenate([sk[i], sk[(i * 7) % 64][:200], rng.integers(0, 2**64 - 4, 900, dtype=np.uint64)]))
    radio = tc.radio(0.05, 21)
    keys = {}
    for name, env in (("on", {}), ("off", {"TCLUST_TILE_FLAGS": "0"}), ("join", {"TCLUST_PASS_IMPL": "join"})):
        for kk, vv in env.items():
            monkeypatch.setenv(kk, vv)
        with tc.Context(0) as c:
            c.upload(sk)
            c.allpairs_pass(radio)
            e = c.edges()
        for kk in env:
            monkeypatch.delenv(kk)
        keys[name] = np.sort(e["i"].astype(np.uint64) << np.uint64(40) | e["j"].astype(np.uint64) << np.uint64(16) | e["common"].astype(np.uint64))
    assert np.array_equal(keys["on"], keys["off"])
    assert np.array_equal(keys["on"], keys["join"])


def _group_mst(tc, sk, k, thr, containment, nranks, devices, phased=False):
    """nranks contexts (threads of this process) form a group and build the MST collectively.  phased: dense pass,
    host barrier, then the collective Boruvka alone (tc_mst_of_edges) -- needed when the members SHARE a device: a
    member's pass kernel (all registers of every SM) cannot start while another member's persistent Boruvka kernel
    waits for it at the group barrier.  With one GPU per member (the real configuration) mst() is used as is."""
    ctxs = [tc.Context(d) for d in devices]
    out, err = [None] * nranks, [None] * nranks
    bar = threading.Barrier(nranks)
    try:
        handles = [c.mg_export(len(sk)) for c in ctxs]

        def work(r):
            try:
                ctxs[r].mg_join(r, nranks, handles)
                ctxs[r].upload(sk)
                for _ in range(2):                             # two collective calls on the same group
                    if phased:
                        ctxs[r].allpairs_pass(tc.radio(thr, k), r, nranks)
                        bar.wait()
                        out[r] = ctxs[r].mst_of_edges(k, int(containment))
                        bar.wait()
                    else:
                        out[r] = ctxs[r].mst(k, thr, containment)
            except Exception as ex:      # noqa: BLE001
                bar.abort()
                err[r] = ex
        th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
        for t in th:
            t.start()
        for t in th:
            t.join()
    finally:
        for c in ctxs:
            try:
                c.mg_leave()
            finally:
                c.close()
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("containment", [False, True])
def test_group_of_two_on_one_gpu(tc, oracle, containment, monkeypatch):
    """The multi-GPU MST (units dealt round-robin, min-reduction over mapped peer buffers inside the Boruvka kernel)
    with both members on device 0: one CTA per SM each, so the two persistent kernels are co-resident."""
    monkeypatch.setenv("TCLUST_K4_CTAS_PER_SM", "1")
    sk = synth.family_sketches(9, 11, s=600, seed=13) + _graphy_sketches(7, n=120)
    k, thr = 21, 0.05
    res = _group_mst(tc, sk, k, thr, containment, 2, [0, 0], phased=True)
    want = oracle.mst(sk, k, thr, containment)
    _same_forest(oracle, res[0], want, len(sk), thr)
    assert np.array_equal(res[0], res[1]), "the ranks of a group must return the same MST"
    with tc.Context(0) as c:
        c.upload(sk)
        single = c.mst(k, thr, containment)
    assert np.array_equal(single, res[0]), "group result differs from the single-GPU result"


def test_group_on_all_gpus(tc, oracle):
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    g = min(g, 8)
    h, off = synth.family_sketches_fast(50, 40, s=500, seed=17)
    sk = [h[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    res = _group_mst(tc, sk, 21, 0.05, False, g, list(range(g)))
    with tc.Context(0) as c:
        c.upload(sk)
        single = c.mst(21, 0.05)
    for r in range(g):
        assert np.array_equal(single, res[r]), r


def _labels_partition(lab):
    out = {}
    for v, l in enumerate(np.asarray(lab).tolist()):
        out.setdefault(l, []).append(v)
    return sorted(out.values())


@pytest.mark.parametrize("containment", [False, True])
def test_cluster_threshold_equals_mst_cut(tc, oracle, containment):
    """tc_cluster_threshold (union-find epilogue of the dense pass, no edges, no MST) = the partition the reference cuts
    out of its MST at the threshold; labels are smallest ids; parts of a sharded run merge to the same partition."""
    sk = synth.family_sketches(7, 15, s=500, seed=31) + _graphy_sketches(5, n=150)
    n = len(sk)
    for thr in (0.02, 0.05, 0.3):
        ref = oracle.mst(sk, 21, thr, containment)
        _, lab_ref, _ = oracle.clusters_from_mst(ref, n, thr)
        with tc.Context(0) as c:
            c.upload(sk)
            nc, lab = c.cluster_threshold(21, thr, containment)
            parts = [c.cluster_threshold(21, thr, containment, p, 3)[1] for p in range(3)]
            with pytest.raises(tc.TclustError, match="part of the candidate pairs"):
                c.dense_ani(21, containment)
        assert _labels_partition(lab) == _labels_partition(lab_ref), thr
        assert nc == len(set(lab.tolist()))
        assert all(min(m) == l for l, m in zip(sorted(set(lab.tolist())), sorted(_labels_partition(lab)))), "labels are smallest ids"
        merged = parts[0]
        for p in parts[1:]:
            ncm, merged = tc.labels_merge(merged, p)
        assert np.array_equal(merged, lab) and ncm == nc


@pytest.mark.parametrize("order", ["family", "shuffled"])
def test_bounded_memory_mst(tc, order, monkeypatch):
    """A candidate graph that does not fit the edge buffer (TCLUST_EDGE_CAP): the pass runs in batches, Boruvka folds every
    full buffer into a forest that stays on the device; the MST is the one of the unbounded build, edge for edge."""
    h, off = synth.family_sketches_fast(3, 800, s=200, seed=23)          # three big families: ~9.6e5 candidate pairs
    n = len(off) - 1
    if order == "shuffled":
        perm = np.random.default_rng(9).permutation(n)
        h = np.ascontiguousarray(h.reshape(n, 200)[perm].reshape(-1))
        monkeypatch.setenv("TCLUST_REORDER", "1")                        # run on the reordered copy (edge ids are mapped back per batch)
    with tc.Context(0) as c:
        c.upload(hashes=h, off=off)
        want = c.mst(21, 0.05).copy()
        assert c.mst_stats()["bounded_batches"] == 0 and c.mst_stats()["edges"] > 300000
    monkeypatch.setenv("TCLUST_EDGE_CAP", str(32 * n * 2))               # two whole-row units per batch at most
    with tc.Context(0) as c:
        c.upload(hashes=h, off=off)
        got = c.mst(21, 0.05).copy()
        st = c.mst_stats()
        with pytest.raises(tc.TclustError, match="TCLUST_EDGE_CAP"):
            c.allpairs_pass(tc.radio(0.05, 21))                           # the full candidate list itself cannot be held
    assert st["bounded_batches"] >= 3 and st["bounded_folds"] >= 2, st
    assert np.array_equal(got, want)
    monkeypatch.setenv("TCLUST_EDGE_CAP", "1000")
    with tc.Context(0) as c:
        c.upload(hashes=h, off=off)
        with pytest.raises(tc.TclustError, match="one work unit"):
            c.mst(21, 0.05)


def test_k3s_long_union_lists(tc, oracle):
    """Union lists of ~14,000 entries per block: K3S streams them in several stages on both sides (4096 row entries,
    2048 column entries per stage, advanced like the inputs of a merge).  Every candidate edge against the oracle."""
    rng = np.random.default_rng(77)
    universe = rng.choice(2**40, 30000, replace=False).astype(np.uint64)
    sk = [np.sort(rng.choice(universe, int(rng.integers(300, 900)), replace=False)) for _ in range(200)]
    sk[17] = sk[16].copy()
    n = len(sk)
    want = np.tril(oracle.common_matrix(sk), -1)
    with tc.Context(0) as c:
        c.upload(sk)
        npairs, nedges = c.allpairs_pass(1 << 30)
        e = c.edges()
        st_tiles = None
        got = np.zeros((n, n), np.uint32)
        got[e["i"], e["j"]] = e["common"]
        assert nedges == len(e) == int((want > 0).sum())
        assert np.array_equal(got, want)
        m = c.mst(21, 0.2)
        st_tiles = c.mst_stats()
    ref = oracle.mst(sk, 21, 0.2)
    _same_forest(oracle, m, ref, n, 0.2)
    assert st_tiles["tiles_flagged"] == st_tiles["tiles_total"] == 28


@pytest.mark.parametrize("n", [1, 2, 3, 33])
def test_tiny_sets(tc, oracle, n):
    """One, two, three sketches and one sketch more than a block: MST, threshold clustering and a group of one."""
    rng = np.random.default_rng(n)
    base = np.unique(rng.integers(0, 2**40, 300, dtype=np.uint64))
    sk = [np.sort(rng.choice(base, 200, replace=False)) for _ in range(n)]
    if n >= 3:
        sk[2] = np.unique(rng.integers(2**41, 2**42, 150, dtype=np.uint64))      # shares nothing with the others
    with tc.Context(0) as c:
        c.upload(sk)
        m = c.mst(21, 0.5)
        ref = oracle.mst(sk, 21, 0.5)
        _same_forest(oracle, m, ref, n, 0.5)
        nc, lab = c.cluster_threshold(21, 0.5)
        _, lab_ref, _ = oracle.clusters_from_mst(ref, n, 0.5)
        assert _labels_partition(lab) == _labels_partition(lab_ref)
        h = c.mg_export(n)
        c.mg_join(0, 1, [h])
        assert np.array_equal(c.mst(21, 0.5), m)
        c.mg_leave()
