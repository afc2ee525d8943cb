"""Wall-clock phases of one end-to-end step (what bench.py's e2e leg times): upload, then tc_mst (dense pass + device
Boruvka + sorted MST on the host).  TCLUST_TIMING=1 adds the library's own phase lines of the upload on stderr."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rabbittclust_b200 as T
from rabbittclust_b200 import synth
h, off = synth.family_sketches_fast(1000, 100, 1000, seed=7)
if len(sys.argv) > 1 and sys.argv[1] == "shuffled":
    p = np.random.default_rng(1).permutation(100000)
    h = h.reshape(100000, 1000)[p].reshape(-1).copy()
hp = torch.from_numpy(h).pin_memory().numpy()
ctx = T.Context(0)
for it in range(4):
    t0 = time.perf_counter(); ctx.upload(hashes=hp, off=off)
    t1 = time.perf_counter(); m = ctx.mst(21, 0.05)
    t2 = time.perf_counter(); m = ctx.mst(21, 0.05)
    t3 = time.perf_counter()
    s = ctx.mst_stats()
    print("step %d: upload %.2f  first tc_mst %.2f  second tc_mst %.2f ms (pass kernel %.3f, boruvka kernel %.3f, forest sort %.3f; "
          "%d rounds, %d edges, %d of %d tiles)" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, s["pass_ms"], s["boruvka_ms"],
                                                    s["sort_ms"], s["rounds"], s["edges"], s["tiles_flagged"], s["tiles_total"]))
