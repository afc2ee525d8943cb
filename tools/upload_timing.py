"""Phase times of tc_sketches_upload at the bench size (TCLUST_TIMING=1)."""
import os, sys, time
os.environ["TCLUST_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rabbittclust_b200 as T
from rabbittclust_b200 import synth
h, off = synth.family_sketches_fast(1000, 100, 1000, seed=7)
hp = torch.from_numpy(h).pin_memory().numpy()
ctx = T.Context(0)
for it in range(3):
    sys.stderr.write("--- upload %d\n" % it)
    t0 = time.time(); ctx.upload(hashes=hp, off=off); sys.stderr.write("total %.1f ms\n" % ((time.time() - t0) * 1e3))
