"""world_size-2 gloo test (CPU) of the multi-GPU MST host logic: tile ownership, two-phase all-reduce(MIN)
Boruvka, host distances.  The per-rank CUDA reductions are replaced by a numpy restatement
(tests/_dist_worker.py); the result must equal the oracle's MST of the whole set on every rank."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from rabbittclust_b200 import mst_dist, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("containment", [0, 1])
def test_two_rank_boruvka_matches_oracle(oracle, tmp_path, containment):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), str(r), "2", str(port),
                               str(containment), str(tmp_path)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    if containment:
        sk = synth.family_sketches(6, 15, seed=21, ragged=(100, 1500))
    else:
        sk = synth.family_sketches(7, 14, s=600, seed=20)
    n = len(sk)
    ref = oracle.mst(sk, 21, 0.05, bool(containment))
    m0, m1 = (np.load(tmp_path / ("mst_%d.npy" % r)) for r in range(2))
    assert np.array_equal(m0, m1), "ranks disagree"
    assert len(m0) == len(ref)
    assert np.array_equal(np.sort(m0["dist"]), np.sort(ref["dist"])), "MST weight multiset differs (bit-exact)"
    nc0, lab0, _ = oracle.clusters_from_mst(m0, n, 0.05)
    nc1, lab1, _ = oracle.clusters_from_mst(ref, n, 0.05)
    part = lambda lab: sorted(sorted(np.nonzero(lab == c)[0].tolist()) for c in range(lab.max() + 1))
    assert nc0 == nc1 and part(lab0) == part(lab1)
    p0, p1 = (np.load(tmp_path / ("pairs_%d.npy" % r)) for r in range(2))
    assert p0[0] + p1[0] == n * (n - 1) // 2          # the shards tile the pair space exactly
    assert p0[1] > 0 and p1[1] > 0


def test_owned_pairs_partition():
    for n in (1, 31, 32, 33, 100, 257):
        for world in (1, 2, 3, 8):
            assert sum(mst_dist.owned_pairs(n, r, world) for r in range(world)) == n * (n - 1) // 2


def test_unit_width_rule_and_dealing_at_bench_size():
    # kj_pick_cseg mirror: at the bench size a unit is a whole row of tiles on 1..8 GPUs (2 units per CTA and rank);
    # small sets keep K3H's 4096 columns; the round-robin dealing tiles the pair space whatever the width
    assert mst_dist.join_cseg(100000, 1) == mst_dist.join_cseg(100000, 8) == 1 << 20
    assert len(mst_dist.join_units(100000, nparts=8)) == 3125
    assert mst_dist.join_cseg(500000, 8) == 1 << 20
    assert mst_dist.join_cseg(20000, 1) == 1 << 20 and mst_dist.join_cseg(20000, 3) == 8192
    assert mst_dist.join_cseg(20000, 8) == 4096 and mst_dist.join_cseg(300, 2) == 4096
    for n, world in ((100000, 8), (20000, 4), (20000, 3)):
        assert sum(mst_dist.owned_pairs(n, r, world) for r in range(world)) == n * (n - 1) // 2


def test_shard_bounds_cover_everything():
    from rabbittclust_b200 import mst_dist
    for total in (0, 1, 7, 1000, 100000001):
        for world in (1, 2, 3, 8):
            spans = [mst_dist.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all((hi - lo) % 2 == 0 for lo, hi in spans[:-1])


def test_value_range_splitters_partition_evenly():
    """VrRank.splitters (host side of the value-range sharded upload): G-1 ascending splitters from the gathered samples;
    bucket = number of splitters <= hash is monotone in the hash and balances the sample."""
    from rabbittclust_b200.mst_dist import VrRank, U64_MAX
    rng = np.random.default_rng(5)
    smp = rng.integers(0, 2**63, 5000, dtype=np.uint64)
    smp[::7] = U64_MAX                                  # padding entries of short shares
    for world in (1, 2, 5, 8):
        spl = VrRank.splitters(smp, world)
        assert len(spl) == world - 1 and np.all(np.diff(spl.astype(np.float64)) >= 0)
        real = np.sort(smp[smp != U64_MAX])
        bucket = np.searchsorted(spl, real, side="right")          # = number of splitters <= x
        assert np.all(np.diff(bucket) >= 0)
        counts = np.bincount(bucket, minlength=world)
        assert counts.max() - counts.min() <= 2
    assert np.array_equal(VrRank.splitters(np.full(10, U64_MAX), 4), np.zeros(3, np.uint64))
