"""The C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol that
include/tclust.h declares; compute entry points fail loudly when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tclust.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tc_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(tc):
    assert declared_symbols() == sorted(tc.SYMBOLS)


def test_library_exports_every_declared_symbol(tc):
    from rabbittclust_b200 import build
    build.build()
    L = ctypes.CDLL(tc.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), s
    assert tc.lib().tc_version() == 100


def test_host_distance_helpers_match_oracle(tc, oracle):
    L = oracle.lib()
    for (c, s0, s1, k) in [(581, 1000, 1000, 21), (1000, 1000, 1000, 21), (1, 1000, 997, 17), (350, 1000, 1000, 21),
                           (77, 100, 9000, 19), (100, 100, 9000, 19)]:
        assert tc.mash_dist(c, s0, s1, k) == L.oc_mash_dist_A(c, s0, s1, k)
        assert tc.aaf_dist(c, s0, s1, k) == L.oc_aaf_dist_mst(c, s0, s1, k)
    assert tc.mash_dist(581, 1000, 1000, 21) == 0.02585735819667742      # SURVEY.md 8c known answer
    assert tc.radio(0.05, 21) == 4 == L.oc_radio(0.05, 21)
    assert L.oc_greedy_common_min(0.05, 21, 1000, 1000, 0) == 350           # SURVEY.md 8a-15 known answer


def test_no_silent_cpu_fallback(tc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tc.TclustError):
        tc.Context(0)


def test_premsted_tree_outputs_match_reference_golden(tmp_path):
    """clust-mst --premsted needs no GPU (host-only consumers of the saved tree): cluster file, Newick /
    PHYLIP / NEXUS trees and the linkage matrix are byte-identical to the reference's own run on the same folder
    (tests/golden/trees = `oracle/_ref/clust-mst -t 1 --premsted tests/golden/mst_list -d 0.05 -o mst.txt
    --newick-tree --phylip-tree --nexus-tree --linkage-matrix`)."""
    import subprocess
    host = os.path.join(ROOT, "rabbittclust_b200", "host")
    from rabbittclust_b200 import build
    build.build()
    subprocess.check_call(["make", "-s", "-C", host])
    gold = os.path.join(ROOT, "tests", "golden")
    r = subprocess.run([os.path.join(host, "bin", "clust-mst"), "--premsted", os.path.join(gold, "mst_list"), "-d", "0.05", "-o",
                        "mst.txt", "--newick-tree", "--phylip-tree", "--nexus-tree", "--linkage-matrix"], cwd=tmp_path,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-1000:]
    for suffix in ("", ".newick.tree", ".phylip.tree", ".nexus.tree", ".linkage.txt"):
        assert open(os.path.join(gold, "trees", "mst.txt" + suffix)).read() == open(tmp_path / ("mst.txt" + suffix)).read(), suffix


def test_labels_merge_host_helper(tc):
    """tc_labels_merge (host side of the sharded threshold-only clustering): union of partial partitions given as
    smallest-id labels, checked against a plain union-find."""
    import numpy as np
    rng = np.random.default_rng(3)
    n = 500

    def random_partition():
        p = np.arange(n)
        for a, b in rng.integers(0, n, (180, 2)):
            ra, rb = a, b
            while p[ra] != ra:
                ra = p[ra]
            while p[rb] != rb:
                rb = p[rb]
            if ra != rb:
                p[max(ra, rb)] = min(ra, rb)
        lab = np.array([v for v in range(n)])
        for v in range(n):
            r = v
            while p[r] != r:
                r = p[r]
            lab[v] = r
        return lab.astype(np.int32)

    a, b = random_partition(), random_partition()
    nc, m = tc.labels_merge(a, b)
    # reference: components of the graph {v - a[v]} u {v - b[v]}
    p = np.arange(n)
    def find(x):
        while p[x] != x:
            x = p[x]
        return x
    for lab in (a, b):
        for v in range(n):
            ra, rb = find(v), find(int(lab[v]))
            if ra != rb:
                p[max(ra, rb)] = min(ra, rb)
    want = np.array([find(v) for v in range(n)], np.int32)
    assert np.array_equal(m, want) and nc == len(set(want.tolist()))
    with pytest.raises(tc.TclustError):
        tc.labels_merge(np.array([0, 5], np.int32), np.array([0, 0], np.int32))     # label out of range
