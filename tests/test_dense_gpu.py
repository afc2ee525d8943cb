"""GPU parity of the --dense epilogue (K6): density maps and ANI histogram over the candidate pairs,
against the reference's own mst.ani / mst.dense (tests/golden/dense, tools/make_golden_dense.py), the oracle
restatement, and -- through the C++ host binary -- the reference's noise-removed cluster files."""
import glob
import os
import subprocess

import numpy as np
import pytest

from rabbittclust_b200 import sketch_io, synth
from test_oracle_golden import _read_dense, partition

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
OURS = os.path.join(ROOT, "rabbittclust_b200", "host", "bin", "clust-mst")


@pytest.mark.parametrize("tag,contain", [("mash", False), ("contain", True)])
def test_dense_ani_matches_reference_files(ctx, tag, contain):
    d = os.path.join(GOLD, "dense", tag)
    hs = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    ctx.upload(hs["sketches"])
    mst = ctx.mst(hs["k"], 0.05, contain)                     # runs the dense pass, keeps the candidate edges
    dense, ani = ctx.dense_ani(hs["k"], contain)
    assert np.array_equal(ani, np.fromfile(os.path.join(d, "mst.ani"), "<u8"))
    assert np.array_equal(dense, _read_dense(os.path.join(d, "mst.dense")))
    ref = sketch_io.read_edge_mst(os.path.join(d, "edge.mst"))
    assert np.array_equal(np.sort(mst["dist"]), np.sort(ref["dist"]))


@pytest.mark.parametrize("contain", [False, True])
def test_dense_ani_vs_oracle_random(ctx, oracle, contain):
    # ragged sizes, identical sketches (J = 1 -> distance 0, ANI bin 100), pairs removed by the size-ratio filter
    sk = synth.family_sketches(6, 12, s=600, seed=41, keep_lo=0.05, ragged=(40, 600))
    sk += synth.family_sketches(3, 10, s=300, seed=42, keep_lo=0.8)
    sk.append(sk[3].copy())
    sk.append(sk[80].copy())
    n = len(sk)
    for k, thr, span in ((21, 0.05, 100), (15, 0.2, 37)):
        ctx.upload(sk)
        from rabbittclust_b200 import radio
        ctx.allpairs_pass(radio(thr, k))
        dense, ani = ctx.dense_ani(k, contain, span)
        wd, wa = oracle.dense_ani(sk, k, thr, contain, span)
        assert np.array_equal(ani, wa), (k, contain)
        assert np.array_equal(dense, wd), (k, contain)
        assert int(ani.sum()) == len(ctx.edges())
        assert ani[100] >= 2


def _run(args, cwd):
    os.makedirs(cwd, exist_ok=True)
    r = subprocess.run([OURS] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    folders = [p for p in glob.glob(os.path.join(cwd, "20*")) if os.path.isdir(p)]
    return folders[0] if folders else None


@pytest.mark.parametrize("tag,extra", [("mash", ["-s", "1000"]), ("contain", ["-c", "100"])])
def test_cli_dense_golden(tmp_path, tag, extra):
    if not os.path.exists(OURS):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "rabbittclust_b200", "host")])
    os.symlink(os.path.join(ROOT, "tests"), str(tmp_path / "tests"))
    of = _run(["--dense", "-l", "-i", "tests/golden/mst_genomes/list.txt", "-k", "15", "-d", "0.05", "-o", "clusters.txt"] + extra,
              str(tmp_path))
    d = os.path.join(GOLD, "dense", tag)
    for fn in ("hash.sketch", "mst.ani", "mst.dense"):
        assert open(os.path.join(d, fn), "rb").read() == open(os.path.join(of, fn), "rb").read(), fn
    for fn in ("clusters.txt", "clusters.txt.removeNoise"):
        assert partition(sketch_io.parse_cluster_file(os.path.join(d, fn))) == \
               partition(sketch_io.parse_cluster_file(str(tmp_path / fn))), fn
    # --presketched --dense: noise genomes of the containment set
    _run(["--dense", "--presketched", os.path.join(GOLD, "presketched", tag), "-d", "0.05", "-o", "p.txt"], str(tmp_path))
    assert partition(sketch_io.parse_cluster_file(os.path.join(GOLD, "dense", "presk_%s.removeNoise" % tag))) == \
           partition(sketch_io.parse_cluster_file(str(tmp_path / "p.txt.removeNoise")))
