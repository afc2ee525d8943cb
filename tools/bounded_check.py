"""Bounded-memory MST at full size: a one-species set -- every pair is a candidate -- built once with
the candidate graph in memory and once with the edge buffer capped (TCLUST_EDGE_CAP); the two MSTs must be identical.
usage: python tools/bounded_check.py [n] [edge_cap_entries]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rabbittclust_b200 as T
from rabbittclust_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 27
h, off = synth.family_sketches_fast(1, n, 1000, seed=5)
res = {"n": n, "candidate_pairs": n * (n - 1) // 2}
out = {}
for name, env in (("in_memory", None), ("bounded", str(cap))):
    if env:
        os.environ["TCLUST_EDGE_CAP"] = env
    with T.Context(0) as c:
        c.upload(hashes=h, off=off)
        c.mst(21, 0.05)
        t0 = time.time(); m = c.mst(21, 0.05).copy(); dt = time.time() - t0
        st = c.mst_stats()
    out[name] = m
    res[name] = {"mst_ms": round(dt * 1e3, 2), "mst_edges": int(len(m)), "batches": st["bounded_batches"], "forest_reductions": st["bounded_folds"] + (1 if st["bounded_batches"] else 0),
                 "edge_buffer_entries": cap if env else None, "edge_buffer_gb": round(cap * 12 / 1e9, 2) if env else round(n * (n - 1) // 2 * 12 / 1e9, 2)}
res["identical"] = bool(np.array_equal(out["in_memory"], out["bounded"]))
print(json.dumps(res))
assert res["identical"]
