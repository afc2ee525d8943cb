"""Quick timing of upload + tc_mst on the bench set, the shuffled set and a one-species set (development aid).
usage: python tools/quick_mst.py [family|shuffled|species] [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rabbittclust_b200 as T
from rabbittclust_b200 import synth
kind = sys.argv[1] if len(sys.argv) > 1 else "family"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (16384 if kind == "species" else 100000)
if kind == "species":
    h, off = synth.family_sketches_fast(1, n, 1000, seed=5)
else:
    h, off = synth.family_sketches_fast(n // 100, 100, 1000, seed=7)
    if kind == "shuffled":
        p = np.random.default_rng(11).permutation(n)
        h = np.ascontiguousarray(h.reshape(n, 1000)[p].reshape(-1))
hp = torch.from_numpy(h).pin_memory().numpy()
with T.Context(0) as ctx:
    for it in range(3):
        t0 = time.perf_counter(); ctx.upload(hashes=hp, off=off)
        t1 = time.perf_counter(); m = ctx.mst(21, 0.05)
        t2 = time.perf_counter(); m = ctx.mst(21, 0.05)
        t3 = time.perf_counter()
        s = ctx.mst_stats()
        print("%s n=%d it%d: upload %.2f  tc_mst %.2f / %.2f ms  pass %.3f  boruvka %.3f  sort %.3f  rounds %d edges %d tiles %d/%d shadow %d mst %d" % (
            kind, n, it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, s["pass_ms"], s["boruvka_ms"], s["sort_ms"], s["rounds"], s["edges"],
            s["tiles_flagged"], s["tiles_total"], s["on_reordered_copy"], len(m)))
