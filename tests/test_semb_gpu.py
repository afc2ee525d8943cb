"""GPU parity of the semantic-B (brute-force) paths -- SURVEY 8a rows a-11, a-12, a-16:
Sketch::MinHash::jaccard/distance/containDistance (RabbitSketch/src/MinHash.cpp:908-1153), modifyMST
(src/MST.cpp:809-1018) and greedyCluster (src/greedy.cpp:285-351) -- against the golden vectors produced by the
unmodified reference (tests/golden/semb, tools/make_golden_semb.py), the oracle restatement and, when oracle/_ref
is present, the reference harness run live."""
import os
import struct
import subprocess

import numpy as np
import pytest

from rabbittclust_b200 import sketch_io, synth
from test_oracle_golden import SEMB_CASES, semb_case

pytestmark = pytest.mark.gpu


def _want_common_b(oracle, sk, s):
    return oracle.common_B_matrix(sk, s)


@pytest.mark.parametrize("name", SEMB_CASES)
def test_pair_common_b_and_distances_golden(ctx, tc, oracle, name):
    z, sk, s, k, thr = semb_case(name)
    n = len(sk)
    ctx.upload(sk)
    got = ctx.pair_common_b(0, n, 0, n, s)
    assert np.array_equal(got, _want_common_b(oracle, sk, s))
    full = ctx.pair_common(0, n, 0, n, impl=2)
    pairs = z[name + "_pairs"]
    idx = 0
    for i in range(n):
        for j in range(i):
            si, sj = len(sk[i]), len(sk[j])
            jac = float(tc.lib().tc_jaccard_b(int(got[i, j]), si, sj, s))
            assert jac == pairs[idx, 0]
            assert tc.distance_b(int(got[i, j]), si, sj, s, k) == pairs[idx, 1]
            assert tc.contain_distance(int(full[i, j]), si, sj, k) == pairs[idx, 2]
            idx += 1


@pytest.mark.parametrize("name", SEMB_CASES)
def test_mst_and_greedy_bruteforce_golden(ctx, name):
    z, sk, s, k, thr = semb_case(name)
    n = len(sk)
    ctx.upload(sk)
    for tag, contain in (("", False), ("_contain", True)):
        mst = ctx.mst_bruteforce(k, s, contain)
        assert len(mst) == n - 1
        assert np.array_equal(np.sort(mst["dist"]), z[name + "_mst_w" + tag]), tag
        assert (mst["pre"] > mst["suf"]).all()
        nc, rep = ctx.greedy_cluster_bruteforce(k, s, thr, contain)
        assert np.array_equal(rep, z[name + "_rep" + tag]), tag
        assert nc == len(set(rep.tolist()))


def test_common_b_multichunk_and_small_budget(ctx, oracle):
    """sketches longer than one 888-row chunk, stop sizes smaller/larger than the sketches, subranges"""
    rng = np.random.default_rng(17)
    sk = synth.family_sketches(3, 11, s=2500, seed=21, keep_lo=0.1)
    for t in range(20):
        size = int(np.exp(rng.uniform(np.log(30), np.log(4000))))
        hi = 2**64 - 1 if t % 3 else 2**36
        sk.append(np.unique(rng.integers(0, hi, size, dtype=np.uint64)))
    sk[4] = np.zeros(0, np.uint64)
    sk[5] = sk[6].copy()
    n = len(sk)
    ctx.upload(sk)
    for s in (1, 7, 100, 888, 889, 1000, 2500, 5000, 100000):
        got = ctx.pair_common_b(0, n, 0, n, s)
        assert np.array_equal(got, oracle.common_B_matrix(sk, s)), s
    got = ctx.pair_common_b(3, 40, 11, 29, 1000)
    assert np.array_equal(got, oracle.common_B_matrix(sk, 1000)[3:40, 11:29])
    # a stop size beyond every union reduces semantic B to semantic A
    assert np.array_equal(ctx.pair_common_b(0, n, 0, n, 1 << 30), ctx.pair_common(0, n, 0, n, impl=1))


def test_pass_b_edges_match_matrix(ctx, oracle):
    sk = synth.family_sketches(7, 13, s=300, seed=33, keep_lo=0.05)
    n = len(sk)
    ctx.upload(sk)
    npairs, nedges = ctx.allpairs_pass_b(300)
    assert npairs == n * (n - 1) // 2
    e = ctx.edges()
    want = oracle.common_B_matrix(sk, 300)
    got = np.zeros((n, n), np.uint32)
    got[e["i"], e["j"]] = e["common"]
    assert (e["j"] < e["i"]).all()
    assert np.array_equal(got, np.tril(want, -1))
    # two parts partition the edges
    e0 = (ctx.allpairs_pass_b(300, 0, 2), ctx.edges())[1]
    e1 = (ctx.allpairs_pass_b(300, 1, 2), ctx.edges())[1]
    both = np.sort(np.concatenate([e0, e1]), order=["i", "j"])
    assert np.array_equal(both, np.sort(e, order=["i", "j"]))


def test_bruteforce_live_against_reference(ctx, oracle, tmp_path):
    """600 sketches: modifyMST weights and greedyCluster centres of the reference binary run here"""
    exe = oracle.ref_bin("ref_harness")
    if exe is None:
        pytest.skip("oracle/_ref not built")
    sk = synth.family_sketches(30, 20, s=400, seed=44, keep_lo=0.3)
    n = len(sk)
    sketch_io.write_minhash_folder(str(tmp_path), sk, ["g%d" % i for i in range(n)], k=21, sketch_size=400)
    hs = str(tmp_path / "hash.sketch")
    ctx.upload(sk)
    for contain in (False, True):
        extra = ["contain"] if contain else []
        o = subprocess.check_output([exe, "mstb", hs, str(n), str(os.cpu_count() or 1)] + extra,
                                    stderr=subprocess.DEVNULL)
        m = struct.unpack_from("<Q", o, 0)[0]
        ref = np.frombuffer(o, sketch_io.EDGE_DTYPE, m, 8)
        mst = ctx.mst_bruteforce(21, 400, contain)
        assert np.array_equal(np.sort(mst["dist"]), np.sort(ref["dist"]))
        o = subprocess.check_output([exe, "greedyb", hs, str(n), "1", "0.05"] + extra, stderr=subprocess.DEVNULL)
        want = np.full(n, -1, np.int32)
        for line in o.decode().strip().split("\n"):
            c = [int(x) for x in line.split()]
            want[c] = c[0]
        _, rep = ctx.greedy_cluster_bruteforce(21, 400, 0.05, contain)
        assert np.array_equal(rep, want)
