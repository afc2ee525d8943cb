"""Quick wall-clock probe of upload + one dense pass (development aid, not the bench)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import rabbittclust_b200 as T
from rabbittclust_b200 import synth
c = T.Context(0)
for N in [int(x) for x in (sys.argv[1:] or ["4096", "16384"])]:
    h, off = synth.family_sketches_fast(max(N // 64, 1), 64, 1000, seed=2)
    t = time.time(); c.upload(hashes=h, off=off); t1 = time.time() - t
    t = time.time(); npairs, ne = c.allpairs_pass(4); t2 = time.time() - t
    print(N, "upload %.3fs pass %.3fs pairs/s %.3e edges %d" % (t1, t2, npairs / t2, ne), flush=True)
