"""GPU parity of the all-pairs intersection (K3), the dense pass and the MST build against the
CPU oracle (oracle/tclust_oracle.c, pinned to the reference)."""
import numpy as np
import pytest

from rabbittclust_b200 import synth

pytestmark = pytest.mark.gpu


def _check_matrix(ctx, oracle, sketches):
    ctx.upload(sketches)
    n = len(sketches)
    want = oracle.common_matrix(sketches)
    got_ref = ctx.pair_common(0, n, 0, n, impl=1)
    got = ctx.pair_common(0, n, 0, n, impl=0)
    got_join = ctx.pair_common(0, n, 0, n, impl=2)
    got_block = ctx.pair_common(0, n, 0, n, impl=3)
    assert np.array_equal(got_ref, want), "u64 global-merge kernel differs from oracle"
    assert np.array_equal(got, want), "tiled merge kernel differs from oracle"
    assert np.array_equal(got_join, want), "hash-join kernel differs from oracle"
    assert np.array_equal(got_block, want), "block-join kernel differs from oracle"


def test_common_fixed_size(ctx, oracle):
    _check_matrix(ctx, oracle, synth.family_sketches(6, 17, s=1000, seed=1))


def test_common_small_and_odd_sizes(ctx, oracle):
    rng = np.random.default_rng(3)
    sk = [np.unique(rng.integers(0, 5000, int(sz), dtype=np.uint64)) for sz in rng.integers(0, 300, 70)]
    sk[5] = np.zeros(0, np.uint64)             # empty sketch
    sk[6] = sk[7].copy()                        # identical pair
    sk[8] = np.array([2**64 - 1, ], np.uint64)  # extreme value
    sk[9] = np.array([0, 2**64 - 1], np.uint64)
    _check_matrix(ctx, oracle, sk)


def test_common_ragged_multichunk(ctx, oracle):
    # sizes 100..6000: several value chunks per tile, very different value ranges per sketch
    rng = np.random.default_rng(5)
    sk = []
    for t in range(40):
        size = int(np.exp(rng.uniform(np.log(100), np.log(6000))))
        hi = 2**64 - 1 if t % 3 else 2**40           # some sketches live in a tiny value range
        sk.append(np.unique(rng.integers(0, hi, size, dtype=np.uint64)))
    sk += synth.family_sketches(2, 6, s=3000, seed=9)
    _check_matrix(ctx, oracle, sk)


def test_common_subrange(ctx, oracle):
    sk = synth.family_sketches(4, 25, s=500, seed=11)
    ctx.upload(sk)
    want = oracle.common_matrix(sk)
    for impl in (0, 2, 3):
        got = ctx.pair_common(37, 91, 3, 66, impl=impl)
        assert np.array_equal(got, want[37:91, 3:66]), impl


def test_join_many_segments_and_partitions(ctx):
    """9000 sketches: three column segments per row block; ragged sizes up to 20000 give row blocks with
    dozens of table partitions.  Hash-join kernel against the independent u64 kernel."""
    rng = np.random.default_rng(12)
    h, off = synth.family_sketches_fast(180, 50, s=300, seed=6)
    sk = [h[int(off[i]):int(off[i + 1])] for i in range(9000)]
    for i in rng.integers(0, 9000, 40):
        sk[i] = np.unique(np.concatenate([sk[i], rng.integers(0, 2**64 - 4, int(rng.integers(2000, 20000)), dtype=np.uint64)]))
    ctx.upload(sk)
    for (i0, i1, j0, j1) in [(0, 64, 0, 9000), (4000, 4100, 3000, 9000), (8950, 9000, 0, 9000)]:
        a = ctx.pair_common(i0, i1, j0, j1, impl=2)
        b = ctx.pair_common(i0, i1, j0, j1, impl=1)
        assert np.array_equal(a, b), (i0, i1, j0, j1)


@pytest.mark.parametrize("pass_impl", ["default", "blockjoin", "join", "merge"])
@pytest.mark.parametrize("containment", [False, True])
def test_dense_pass_edges_and_mst(ctx, oracle, tc, containment, pass_impl, monkeypatch):
    if pass_impl != "default":      # "default" = the production routing (K3J per unit where it pays, else K3H; tile flags)
        monkeypatch.setenv("TCLUST_PASS_IMPL", pass_impl)
    k, thr = 21, 0.05
    if containment:
        sk = synth.family_sketches(8, 12, seed=21, ragged=(100, 2000))
    else:
        sk = synth.family_sketches(10, 13, s=1000, seed=20)
    n = len(sk)
    ctx.upload(sk)
    radio = tc.radio(thr, k)
    assert radio == oracle.lib().oc_radio(thr, k)
    npairs, nedges = ctx.allpairs_pass(radio)
    assert npairs == n * (n - 1) // 2
    e = ctx.edges()
    assert len(e) == nedges
    want = oracle.common_matrix(sk)
    sizes = np.array([len(s) for s in sk])
    exp = {}
    for i in range(n):
        for j in range(i):
            c = int(want[i, j])
            mn, mx = min(sizes[i], sizes[j]), max(sizes[i], sizes[j])
            if c >= 1 and mn > 0 and not mx > radio * mn:
                exp[(i, j)] = c
    got = {(int(r["i"]), int(r["j"])): int(r["common"]) for r in e}
    assert got == exp
    mst = ctx.mst(k, thr, containment)
    ref = oracle.mst(sk, k, thr, containment)
    assert len(mst) == len(ref)
    assert np.array_equal(np.sort(mst["dist"]), np.sort(ref["dist"])), "MST weight multiset differs (bit-exact)"
    nc_g, lab_g, _ = oracle.clusters_from_mst(mst, n, thr)
    nc_r, lab_r, _ = oracle.clusters_from_mst(ref, n, thr)
    assert nc_g == nc_r
    part = lambda lab: sorted(sorted(np.nonzero(lab == c)[0].tolist()) for c in range(lab.max() + 1))
    assert part(lab_g) == part(lab_r)


def test_larger_block_against_u64_kernel(ctx):
    # 1500 sketches: the tiled kernel against the independent u64 kernel on a full matrix
    h, off = synth.family_sketches_fast(30, 50, s=1000, seed=4)
    ctx.upload(hashes=h, off=off)
    a = ctx.pair_common(0, 1500, 0, 1500, impl=0)
    b = ctx.pair_common(0, 1500, 0, 1500, impl=1)
    c = ctx.pair_common(0, 1500, 0, 1500, impl=2)
    assert np.array_equal(a, b)
    assert np.array_equal(c, b)
    assert (np.diag(a) == 1000).all()
    assert a[0, 1] >= 250 and a[0, 60] == 0


@pytest.mark.parametrize("containment", [False, True])
def test_sharded_pass_two_parts_equals_oracle(tc, oracle, containment):
    """The multi-GPU split on one GPU: two contexts own the tiles of part 0 and part 1; the element-wise
    minimum of their Boruvka reductions (what the NCCL all-reduce(MIN) computes) yields the reference MST."""
    from rabbittclust_b200 import mst_dist
    k, thr = 21, 0.05
    sk = (synth.family_sketches(8, 12, seed=31, ragged=(100, 2000)) if containment
          else synth.family_sketches(9, 15, s=800, seed=30))
    n = len(sk)
    radio = tc.radio(thr, k)
    ctxs = [tc.Context(0), tc.Context(0)]
    try:
        tot_pairs = 0
        for part, c in enumerate(ctxs):
            c.upload(sk)
            npairs, _ = c.allpairs_pass(radio, part, 2)
            assert npairs == mst_dist.owned_pairs(n, part, 2, impl="join")
            tot_pairs += npairs
        assert tot_pairs == n * (n - 1) // 2
        ops = [mst_dist.GpuOps(c, containment) for c in ctxs]

        class Both:
            def phase1(self, comp):
                return np.minimum(ops[0].phase1(comp), ops[1].phase1(comp))

            def phase2(self, comp, key):
                return np.minimum(ops[0].phase2(comp, key), ops[1].phase2(comp, key))

        L = tc.lib()
        mst = mst_dist.boruvka(n, Both(), lambda key: float(L.tc_dist_from_key(key, int(containment), k)))
    finally:
        for c in ctxs:
            c.close()
    ref = oracle.mst(sk, k, thr, containment)
    assert len(mst) == len(ref)
    assert np.array_equal(np.sort(mst["dist"]), np.sort(ref["dist"]))
    single = tc.Context(0)
    try:
        single.upload(sk)
        one = mst_dist.distributed_mst(single, k, thr, containment)
        assert np.array_equal(np.sort(one["dist"]), np.sort(ref["dist"]))
    finally:
        single.close()


def test_full_size_pass_properties(ctx):
    """BASELINE configs[2] at full size (N = 100,000, s = 1000; 5.0e9 pairs): the oracle cannot run this, so the dense
    pass is checked through properties of the generator -- pool hashes are even and private ones odd, hence a pair
    shares hashes iff it lies inside a family: exactly 1000 * C(100,2) candidate edges, every edge in-family, and
    `common` equal to the exact intersection recomputed on the host for a sample of edges."""
    h, off = synth.family_sketches_fast(1000, 100, 1000, seed=7)
    n = len(off) - 1
    ctx.upload(hashes=h, off=off)
    npairs, nedges = ctx.allpairs_pass(4)
    assert npairs == n * (n - 1) // 2
    assert nedges == 1000 * (100 * 99 // 2)
    e = ctx.edges()
    assert (e["j"] < e["i"]).all() and (e["i"] // 100 == e["j"] // 100).all()
    rows = h.reshape(n, 1000)
    rng = np.random.default_rng(0)
    for t in rng.integers(0, len(e), 300):
        i, j, c = int(e["i"][t]), int(e["j"][t]), int(e["common"][t])
        assert c == len(np.intersect1d(rows[i], rows[j], assume_unique=True))
    # idempotence: a second pass yields the same edge multiset
    ctx.allpairs_pass(4)
    e2 = ctx.edges()
    assert np.array_equal(np.sort(e, order=["i", "j"]), np.sort(e2, order=["i", "j"]))


def test_pass_from_and_mst_append(ctx, oracle):
    """start_index form of the dense pass (compute_minhash_mst with start_index, MST.cpp:1375-1383) and the incremental
    MST: appending the last third to the MST of the first two thirds gives the weights of the MST of everything."""
    from rabbittclust_b200 import radio
    sk = synth.family_sketches(9, 11, s=400, seed=77, keep_lo=0.3)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(sk))
    sk = [sk[i] for i in perm]
    n, start = len(sk), 61                                   # not a multiple of 32
    ctx.upload(sk)
    npairs, nedges = ctx.allpairs_pass_from(radio(0.05, 21), start)
    assert npairs == n * (n - 1) // 2 - start * (start - 1) // 2
    e = ctx.edges()
    want = oracle.common_matrix(sk)
    got = np.zeros((n, n), np.uint32)
    got[e["i"], e["j"]] = e["common"]
    ref = np.tril(want, -1)
    ref[:start] = 0
    assert np.array_equal(got, ref)
    full = oracle.mst(sk, 21, 0.05)
    ctx.upload(sk[:start])
    pre = ctx.mst(21, 0.05)
    ctx.upload(sk)
    app = ctx.mst_append(21, 0.05, start, pre)
    assert np.array_equal(np.sort(app["dist"]), np.sort(full["dist"]))
    _, la, _ = oracle.clusters_from_mst(app, n, 0.05)
    _, lb, _ = oracle.clusters_from_mst(full, n, 0.05)
    assert np.array_equal(la, lb)
    # start_index = 0 is the ordinary MST, start_index = n keeps the old tree
    assert np.array_equal(np.sort(ctx.mst_append(21, 0.05, 0, pre[:0])["dist"]), np.sort(full["dist"]))


def test_edge_buffer_grows_on_overflow(tc, oracle, monkeypatch):
    """a candidate graph denser than the first edge buffer: the pass overflows, the kernels keep counting, the buffer
    is re-sized to the exact need and the pass repeated; a hard TCLUST_EDGE_CAP makes the same situation an error"""
    sk = synth.family_sketches(4, 30, s=200, seed=5)
    n = len(sk)
    want = np.tril(oracle.common_matrix(sk), -1)
    monkeypatch.setenv("TCLUST_EDGE_CAP_INIT", "100")
    with tc.Context(0) as c:
        c.upload(sk)
        for impl in ("join", "merge"):
            monkeypatch.setenv("TCLUST_PASS_IMPL", impl)
            _, nedges = c.allpairs_pass(1 << 30)
            assert nedges == int((want > 0).sum()) > 100
            e = c.edges()
            got = np.zeros((n, n), np.uint32)
            got[e["i"], e["j"]] = e["common"]
            assert np.array_equal(got, want)
    monkeypatch.delenv("TCLUST_EDGE_CAP_INIT")
    monkeypatch.delenv("TCLUST_PASS_IMPL")
    monkeypatch.setenv("TCLUST_EDGE_CAP", "100")
    with tc.Context(0) as c:
        c.upload(sk)
        with pytest.raises(tc.TclustError, match="TCLUST_EDGE_CAP"):
            c.allpairs_pass(1 << 30)


def test_piecewise_upload_checks_every_sketch(tc, monkeypatch):
    """the upload in pieces (TCLUST_UPLOAD_OVERLAP=1 forces it for a small set) checks a sketch when the piece holding its
    last hash has landed: an offending sketch anywhere -- first piece, a sketch straddling two pieces, last piece -- is named"""
    monkeypatch.setenv("TCLUST_UPLOAD_OVERLAP", "1")
    rng = np.random.default_rng(3)
    base = [np.unique(rng.integers(0, 1 << 40, int(z), dtype=np.uint64)) for z in rng.integers(1, 3000, 40)]
    base.insert(7, np.zeros(0, np.uint64))                                  # an empty sketch in the middle
    with tc.Context(0) as c:
        c.upload(base)
        want = c.pair_common(0, len(base), 0, len(base), impl=2)
        total = sum(len(b) for b in base)
        for idx in (0, 3, 19, 24, len(base) - 1):
            if len(base[idx]) < 2:
                continue
            bad = [b.copy() for b in base]
            j = len(bad[idx]) // 2
            bad[idx][j] = bad[idx][j - 1]                                    # a duplicate somewhere inside sketch idx
            with pytest.raises(tc.TclustError, match="sketch %d is not strictly ascending" % idx):
                c.upload(bad)
        c.upload(base)                                                      # the context is usable again afterwards
        assert np.array_equal(c.pair_common(0, len(base), 0, len(base), impl=2), want)
        assert total > 8 * 1024                                             # several pieces of >= 1024 hashes


def test_upload_rejects_unsorted_or_duplicate_sketches(tc):
    a = np.array([5, 9, 9, 12], np.uint64)           # duplicate
    b = np.array([3, 2, 7], np.uint64)               # not ascending
    ok = np.array([1, 2, 3], np.uint64)
    with tc.Context(0) as c:
        for bad, idx in (([ok, a], 1), ([b, ok], 0), ([ok, ok, b], 2)):
            with pytest.raises(tc.TclustError, match="sketch %d is not strictly ascending" % idx):
                c.upload(bad)
        c.upload([ok, ok])
        assert c.pair_common(0, 2, 0, 2, impl=2).tolist() == [[3, 3], [3, 3]]
        with pytest.raises(tc.TclustError):
            c.pair_common(0, 3, 0, 2)                 # range out of bounds


def test_randomised_shapes_vs_oracle(ctx, oracle):
    """random mixtures of sizes, value ranges, duplicates of whole sketches and empty sketches: all three kernels
    against the oracle, plus the semantic-B kernel at a random stop size"""
    rng = np.random.default_rng(2024)
    for trial in range(6):
        n = int(rng.integers(2, 90))
        hi = int(rng.choice([300, 5000, 2**20, 2**63]))
        sk = []
        for _ in range(n):
            size = int(rng.choice([0, 1, 2, 31, 32, 33, 200, 887, 888, 889, 1500]))
            size = min(size, hi)
            sk.append(np.sort(rng.choice(hi, size, replace=False).astype(np.uint64)) if hi < 2**21
                      else np.unique(rng.integers(0, hi, size, dtype=np.uint64)))
        for _ in range(3):
            i, j = rng.integers(0, n, 2)
            sk[i] = sk[j].copy()
        ctx.upload(sk)
        want = oracle.common_matrix(sk)
        for impl in (0, 1, 2, 3):
            assert np.array_equal(ctx.pair_common(0, n, 0, n, impl=impl), want), (trial, impl)
        s = int(rng.choice([1, 50, 888, 1000, 4000]))
        assert np.array_equal(ctx.pair_common_b(0, n, 0, n, s), oracle.common_B_matrix(sk, s)), (trial, s)


def test_blockjoin_pass_equals_join_pass(tc, monkeypatch):
    """K3J (block join) dense pass against K3H on 9000 sketches: three column segments, row blocks with 1..3 table
    partitions, ragged blocks without a union list (routed to K3H), plus the start_index form"""
    rng = np.random.default_rng(8)
    sk = synth.family_sketches(30, 100, s=700, seed=61, keep_lo=0.3)                 # related neighbours: P = 1
    perm = rng.permutation(3000)
    sk += [sk[i] for i in perm[:2500]]                                               # unrelated neighbours: P = 2..3
    sk += synth.family_sketches(5, 100, s=1000, seed=62, ragged=(300, 5000))         # blocks above 32768 elements
    sk += [np.unique(rng.integers(0, 2**40, int(z), dtype=np.uint64)) for z in rng.integers(0, 1500, 3000)]
    n = len(sk)
    with tc.Context(0) as c:
        c.upload(sk)
        res = {}
        for impl in ("join", "blockjoin", "auto"):
            if impl == "auto":
                monkeypatch.delenv("TCLUST_PASS_IMPL")            # default: per-unit routing between K3J and K3H
            else:
                monkeypatch.setenv("TCLUST_PASS_IMPL", impl)
            npairs, nedges = c.allpairs_pass(4)
            assert npairs == n * (n - 1) // 2
            res[impl] = np.sort(c.edges(), order=["i", "j"])
            c.allpairs_pass_from(4, 7013)
            res[impl + "_from"] = np.sort(c.edges(), order=["i", "j"])
            p0 = (c.allpairs_pass(4, 0, 2), np.sort(c.edges(), order=["i", "j"]))[1]
            p1 = (c.allpairs_pass(4, 1, 2), np.sort(c.edges(), order=["i", "j"]))[1]
            res[impl + "_parts"] = np.sort(np.concatenate([p0, p1]), order=["i", "j"])
        assert len(res["join"]) > 100000
        assert np.array_equal(res["join"], res["blockjoin"])
        assert np.array_equal(res["join_from"], res["blockjoin_from"])
        assert np.array_equal(res["join"], res["blockjoin_parts"]) and np.array_equal(res["join"], res["join_parts"])
        assert np.array_equal(res["join"], res["auto"]) and np.array_equal(res["join_from"], res["auto_from"])
        assert np.array_equal(res["join"], res["auto_parts"])


def test_reordered_copy_gives_the_same_edges_and_mst(tc, monkeypatch):
    """Dense pass on the reordered shadow copy (k3r_reorder.cu) against the pass in input order: sketches of 40
    families in shuffled order, plus empty and unrelated ones; edge sets (also split over two parts), pair counts and
    the MST must not depend on the internal order"""
    rng = np.random.default_rng(12)
    sk = synth.family_sketches(40, 100, s=600, seed=71, keep_lo=0.3)
    sk += [np.unique(rng.integers(0, 2**40, int(z), dtype=np.uint64)) for z in rng.integers(0, 900, 500)]
    sk += [np.empty(0, np.uint64)] * 7
    order = rng.permutation(len(sk))
    sk = [sk[i] for i in order]
    n = len(sk)
    res = {}
    for mode in ("0", "1", None):
        if mode is None:
            monkeypatch.delenv("TCLUST_REORDER", raising=False)       # default: the upload decides (here: shuffled input, reorder)
        else:
            monkeypatch.setenv("TCLUST_REORDER", mode)
        with tc.Context(0) as c:
            c.upload(sk)
            npairs, nedges = c.allpairs_pass(4)
            assert npairs == n * (n - 1) // 2
            e = np.sort(c.edges(), order=["i", "j"])
            assert nedges == len(e) and np.all(e["j"] < e["i"])
            p = [(c.allpairs_pass(4, q, 3), np.sort(c.edges(), order=["i", "j"])) for q in range(3)]
            assert sum(x[0][0] for x in p) == npairs
            parts = np.sort(np.concatenate([x[1] for x in p]), order=["i", "j"])
            c.allpairs_pass_from(4, n - 333)              # start_index form stays in input order
            frm = np.sort(c.edges(), order=["i", "j"])
            mst = c.mst(21, 0.05, False)
            res[mode] = (e, parts, frm, mst)
    assert len(res["0"][0]) > 100000
    for mode in ("1", None):
        for a, b in zip(res["0"], res[mode]):
            assert np.array_equal(a, b), mode


def test_upload_rows_equals_flat_upload(tc):
    """tc_sketches_upload_rows (one array per sketch, staged in 32 MB chunks) against the flat upload: ragged sizes,
    empty sketches, a set larger than two staging buffers"""
    rng = np.random.default_rng(21)
    sk = synth.family_sketches(12, 20, s=900, seed=81, keep_lo=0.3)
    sk += [np.unique(rng.integers(0, 2**63, int(z), dtype=np.uint64)) for z in rng.integers(0, 3000, 300)]
    sk[7] = np.empty(0, np.uint64)
    sk.append(np.empty(0, np.uint64))
    big = np.unique(rng.integers(0, 2**64 - 2, 9_500_000, dtype=np.uint64))        # 76 MB: spans three chunks
    sk.insert(100, big)
    with tc.Context(0) as a, tc.Context(0) as b:
        a.upload(sk)
        b.upload_rows(sk)
        n = len(sk)
        assert np.array_equal(a.pair_common(0, n, 0, n, impl=1), b.pair_common(0, n, 0, n, impl=1))
        a.allpairs_pass(4); b.allpairs_pass(4)
        assert np.array_equal(np.sort(a.edges(), order=["i", "j"]), np.sort(b.edges(), order=["i", "j"]))
    with tc.Context(0) as c:
        with pytest.raises(tc.TclustError):
            c.upload_rows([np.array([5, 3], np.uint64)])                          # unsorted input is still rejected


def test_chunked_upload_equals_single_copy_upload(tc, oracle, monkeypatch):
    """tc_sketches_upload of a large host array copies it in pieces and sorts/merges the pieces under the copy
    (upload_sorted_runs); TCLUST_UPLOAD_OVERLAP=1 forces that path at any size.  The rank transform it feeds must give the
    same `common` for every pair as the oracle, for totals that leave pieces empty, ragged, or a single element."""
    rng = np.random.default_rng(33)
    sets = [
        synth.family_sketches(5, 9, s=1000, seed=4),                                         # 45 000 hashes: 8 full pieces
        [np.unique(rng.integers(0, 2**64 - 1, int(z), dtype=np.uint64)) for z in rng.integers(0, 120, 40)],  # ~2 400: 3 pieces
        [np.array([7], np.uint64), np.array([7, 9], np.uint64)],                              # 3 hashes: one piece
        [np.unique(rng.integers(0, 3000, int(z), dtype=np.uint64)) for z in rng.integers(500, 2500, 30)],    # many equal hashes
    ]
    monkeypatch.setenv("TCLUST_UPLOAD_OVERLAP", "1")
    with tc.Context(0) as c:
        for sk in sets:
            c.upload(sk)
            n = len(sk)
            want = oracle.common_matrix(sk)
            for impl in (0, 2, 3):
                assert np.array_equal(c.pair_common(0, n, 0, n, impl=impl), want), impl
        with pytest.raises(tc.TclustError, match="sketch 1 is not strictly ascending"):
            c.upload([np.array([1, 2, 3], np.uint64), np.array([5, 5], np.uint64)])
    # and the edges of a dense pass agree with the single-copy upload
    sk = synth.family_sketches(8, 40, s=700, seed=12, keep_lo=0.4)
    with tc.Context(0) as a:
        a.upload(sk); a.allpairs_pass(4); ea = np.sort(a.edges(), order=["i", "j"])
    monkeypatch.setenv("TCLUST_UPLOAD_OVERLAP", "0")
    with tc.Context(0) as b:
        b.upload(sk); b.allpairs_pass(4); eb = np.sort(b.edges(), order=["i", "j"])
    assert np.array_equal(ea, eb)


def test_distinct_sort_blocked_equals_plain(tc, oracle, monkeypatch):
    """k3h_distinct_kernel sorts a block's shared ranks with register-blocked bitonic passes; TCLUST_DISTINCT_SORT=plain
    keeps the one-pass-per-stride version.  Both must give the oracle's counts through the block join (which reads the
    union lists) for short rows (R < 512: mixed plain/blocked stages), long rows and an unsorted-slot block."""
    rng = np.random.default_rng(44)
    sets = [
        synth.family_sketches(3, 40, s=1000, seed=5),                                       # R = 1024
        synth.family_sketches(4, 33, s=200, seed=6, keep_lo=0.2),                            # R = 256
        [np.unique(rng.integers(0, 4000, int(z), dtype=np.uint64)) for z in rng.integers(0, 60, 100)],       # R = 32/64
        [np.unique(rng.integers(0, 4000, int(z), dtype=np.uint64)) for z in ([3000] * 3 + [40] * 61)],       # rows > 1024: slots unsorted
    ]
    for mode in ("blocked", "plain"):
        monkeypatch.setenv("TCLUST_DISTINCT_SORT", mode)
        with tc.Context(0) as c:
            for sk in sets:
                c.upload(sk)
                n = len(sk)
                assert np.array_equal(c.pair_common(0, n, 0, n, impl=3), oracle.common_matrix(sk)), mode


def test_boruvka_async_phases_equal_sync_phases(ctx, tc):
    """tc_boruvka_set_async(1): the phases only enqueue on the library's stream; after one synchronize their buffers
    hold what the synchronous phases return, and tc_mst (which reads every phase's result at once) is unaffected."""
    import ctypes as C
    import torch
    sk = synth.family_sketches(9, 14, s=600, seed=31, keep_lo=0.3)
    n = len(sk)
    ctx.upload(sk)
    ctx.allpairs_pass(4)
    L = tc.lib()
    comp = (np.arange(n, dtype=np.int32) // 3 * 3).astype(np.int32)       # components of three vertices
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    res = {}
    for mode in (0, 1):
        key = torch.empty(n, dtype=torch.int64, device="cuda:0")
        edge = torch.empty(n, dtype=torch.int64, device="cuda:0")
        torch.cuda.synchronize()
        tc._chk(L.tc_boruvka_set_async(ctx._h, mode))
        tc._chk(L.tc_boruvka_upload_comp(ctx._h, comp.ctypes.data_as(C.POINTER(C.c_int32)), n))
        tc._chk(L.tc_boruvka_phase1(ctx._h, 0, key.data_ptr()))
        tc._chk(L.tc_boruvka_phase2(ctx._h, 0, key.data_ptr(), edge.data_ptr()))
        stream.synchronize()
        res[mode] = (key.cpu().numpy(), edge.cpu().numpy())
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert (res[0][1] != -1).any()
    want = ctx.mst(21, 0.05)                                              # async still on: tc_mst forces its own syncs
    tc._chk(L.tc_boruvka_set_async(ctx._h, 0))
    got = ctx.mst(21, 0.05)
    assert np.array_equal(want, got)
