"""Dense-pass kernel time against the units-per-CTA rule of kj_pick_cseg (TCLUST_KJ_MIN_UNITS), whole pass and the
parts of a P-GPU run, on ONE GPU.  usage: python tools/unit_width_sweep.py [N] [P]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rabbittclust_b200 as T
from rabbittclust_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
h, off = synth.family_sketches_fast(n // 100, 100, 1000, seed=7)
for cap, mu in ((1 << 20, 2), (65536, 2), (32768, 2), (1 << 20, 4), (1 << 20, 1)):
    os.environ["TCLUST_KJ_MIN_UNITS"] = str(mu)
    os.environ["TCLUST_KJ_CSEG_MAX"] = str(cap)
    with T.Context(0) as ctx:
        ctx.upload(hashes=h, off=off)
        res = []
        for part, nparts in ((0, 1), (0, 2), (0, 4), (0, P), (P - 1, P)):
            ctx.allpairs_pass(4, part, nparts)
            ks = []
            for _ in range(3):
                ctx.allpairs_pass(4, part, nparts)
                ks.append(ctx.last_pass_kernel_ms)
            res.append("part %d/%d %.3f ms" % (part, nparts, min(ks)))
        print("widest unit %5d, min units per CTA %2d:" % (cap, mu), "  ".join(res), flush=True)
