#!/usr/bin/env python
"""Golden vectors for the semantic-B (brute-force) paths, from the UNMODIFIED reference:
oracle/_ref/ref_harness calls Sketch::MinHash::jaccard/distance/containDistance, modifyMST
(src/MST.cpp:809-1018) and greedyCluster (src/greedy.cpp:285-351) of the reference objects it links.
Writes tests/golden/semb/semb.npz (also run as the last step of tools/make_golden.py).

Per case <c>: <c>_hashes/<c>_off (the sketches), <c>_pairs (rows i, j<i: jaccard, distance, containDistance),
<c>_mst_w / <c>_mst_w_contain (sorted modifyMST weights), <c>_rep / <c>_rep_contain (greedyCluster centres, -t 1).
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabbittclust_b200 import sketch_io, synth  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
CASES = {   # name: (sketches, sketch_size parameter, k, threshold)
    "eq200": (lambda: synth.family_sketches(5, 8, s=200, seed=11), 200, 21, 0.05),
    "rag200": (lambda: synth.family_sketches(5, 8, s=200, seed=12, ragged=(50, 200)), 200, 21, 0.05),
    "rag64": (lambda: synth.family_sketches(4, 9, s=64, seed=13, ragged=(20, 300)), 64, 17, 0.08),
    "eq1000": (lambda: synth.family_sketches(3, 12, s=1000, seed=14, keep_lo=0.2), 1000, 21, 0.05),
}


def main():
    out = {}
    for name, (gen, s, k, thr) in CASES.items():
        sk = gen()
        n = len(sk)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in sk])
        out[name + "_hashes"] = np.concatenate(sk)
        out[name + "_off"] = off
        out[name + "_param"] = np.array([s, k, int(round(thr * 1e6))])
        with tempfile.TemporaryDirectory() as d:
            sketch_io.write_minhash_folder(d, sk, ["g%d" % i for i in range(n)], k=k, sketch_size=s)
            hs = os.path.join(d, "hash.sketch")
            o = subprocess.check_output([EXE, "pairsdump", hs, str(n)])
            out[name + "_pairs"] = np.frombuffer(o, np.float64).reshape(-1, 3).copy()
            for tag, extra in (("", []), ("_contain", ["contain"])):
                o = subprocess.check_output([EXE, "mstb", hs, str(n), "1"] + extra, stderr=subprocess.DEVNULL)
                m = struct.unpack_from("<Q", o, 0)[0]
                e = np.frombuffer(o, sketch_io.EDGE_DTYPE, m, 8)
                out[name + "_mst_w" + tag] = np.sort(e["dist"])
                o = subprocess.check_output([EXE, "greedyb", hs, str(n), "1", repr(thr)] + extra,
                                            stderr=subprocess.DEVNULL).decode()
                rep = np.full(n, -1, np.int32)
                for line in o.strip().split("\n"):
                    c = [int(x) for x in line.split()]
                    rep[c] = c[0]
                out[name + "_rep" + tag] = rep
    dst = os.path.join(ROOT, "tests", "golden", "semb")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "semb.npz"), **out)
    print("wrote", os.path.join(dst, "semb.npz"), os.path.getsize(os.path.join(dst, "semb.npz")), "bytes")


if __name__ == "__main__":
    main()
