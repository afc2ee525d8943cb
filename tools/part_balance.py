"""Time one dense pass restricted to part p of P on ONE GPU (is the multi-GPU shard rule balanced?).
usage: python tools/part_balance.py [N] [P]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rabbittclust_b200 as T
from rabbittclust_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h, off = synth.family_sketches_fast(n // 100, 100, 1000, seed=7)
ctx = T.Context(0)
ctx.upload(hashes=h, off=off)
ctx.allpairs_pass(4)
for rep in range(2):
    t0 = time.time(); ctx.allpairs_pass(4); t_all = time.time() - t0
    parts = []
    for p in range(P):
        ctx.allpairs_pass(4, p, P)          # the unit routing of a part is cached after its first pass
        t0 = time.time(); np_, ne = ctx.allpairs_pass(4, p, P); parts.append((time.time() - t0, ctx.last_pass_kernel_ms, np_, ne))
    print("all %.1f ms" % (t_all * 1e3), " ".join("part%d %.2f ms call / %.2f ms kernel (%d pairs, %d edges)" % (p, t * 1e3, k, a, b) for p, (t, k, a, b) in enumerate(parts)))
