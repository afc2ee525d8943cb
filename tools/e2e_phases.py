"""Wall-clock phases of one end-to-end step (what bench.py's e2e leg times): upload, first pass after it, edge copy."""
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
radio = T.radio(0.05, 21)
for it in range(4):
    t0 = time.perf_counter(); ctx.upload(hashes=hp, off=off)
    t1 = time.perf_counter(); ctx.allpairs_pass(radio)
    t2 = time.perf_counter(); e = ctx.edges(pinned=True)
    t3 = time.perf_counter(); ctx.allpairs_pass(radio)
    t4 = time.perf_counter()
    print("step %d: upload %.2f  first pass %.2f (kernel %.2f)  edges %.2f  second pass %.2f ms" % (
        it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, ctx.last_pass_kernel_ms, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
