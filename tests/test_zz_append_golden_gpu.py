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
