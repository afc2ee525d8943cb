"""Where do the occasional slow tc_allpairs_pass calls come from?  Host clock around the call vs the kernel-only
duration from the library's events (development aid).  usage: [TCLUST_SYNC=spin|yield|block] python tools/step_jitter.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rabbittclust_b200 as T
from rabbittclust_b200 import synth
h, off = synth.family_sketches_fast(1000, 100, 1000, seed=7)
ctx = T.Context(0)
ctx.upload(hashes=h, off=off)
ctx.allpairs_pass(4)
rows = []
for it in range(60):
    t0 = time.perf_counter()
    ctx.allpairs_pass(4)
    t1 = time.perf_counter()
    rows.append(((t1 - t0) * 1e3, ctx.last_pass_kernel_ms))
bad = [(round(a, 1), round(b, 1)) for a, b in rows if a > b + 3]
print(os.environ.get("TCLUST_SYNC", "auto"), "median call ms %.2f" % sorted(a for a, _ in rows)[30], "kernel max %.2f" % max(b for _, b in rows),
      "slow calls:", len(bad), bad[:8])
