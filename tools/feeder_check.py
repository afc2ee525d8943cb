#!/usr/bin/env python
"""Steady-state rate of the FASTA feeder at a size where the CUDA start-up does not dominate: G genomes of
L bp (one FASTA file each, 80 columns) through the drop-in clust-mst and through the unmodified reference at -t nproc;
reports the "computing sketch" phase of both in Gbp/s and checks hash.sketch byte for byte with -t 1 semantics (list order).
usage: python tools/feeder_check.py [workdir] [genomes] [length] [--skip-ref]"""
import glob, json, os, re, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

skip_ref = "--skip-ref" in sys.argv
argv = [a for a in sys.argv if a != "--skip-ref"]
work = argv[1] if len(argv) > 1 else "/tmp/feeder"
G = int(argv[2]) if len(argv) > 2 else 2000
L = int(argv[3]) if len(argv) > 3 else 5000000
sys.path.insert(0, os.path.join(ROOT, "tools"))
from feeder_gen import generate
t0 = time.time()
lst = generate(work, G, L)
res = {"genomes": G, "length": L, "gbp": G * L / 1e9, "generate_s": round(time.time() - t0, 1)}


def run(exe, extra, cwd):
    os.makedirs(cwd, exist_ok=True)
    t = time.time()
    r = subprocess.run([exe] + extra + ["-l", "-i", lst, "-k", "21", "-s", "1000", "-d", "0.05", "-o", "out.txt"], cwd=cwd,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    wall = time.time() - t
    err = r.stderr.decode()
    if r.returncode:
        sys.stderr.write(err[-2000:]); raise SystemExit("failed: " + exe)
    ph = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"=+time of ([^=]+?) is: ([0-9.eE+-]+)", err)}
    folder = [p for p in glob.glob(os.path.join(cwd, "20*")) if os.path.isdir(p)][0]
    return {"wall_s": round(wall, 2), "sketch_s": ph.get("computing sketch"), "gbp_per_s": round(G * L / 1e9 / ph["computing sketch"], 2),
            "timing": [l for l in err.splitlines() if l.startswith("[tclust timing]") and ("context" in l or "feeder" in l)]}, folder

env_t = os.environ.get("TCLUST_TIMING")
os.environ["TCLUST_TIMING"] = "1"
# producer side alone (read + parse on all host threads, nothing handed to the GPU)
exe_ours = os.path.join(ROOT, "rabbittclust_b200", "host", "bin", "clust-mst")
dry = subprocess.run([exe_ours, "-l", "-i", lst, "-o", "o.txt"], cwd=work, env=dict(os.environ, TCLUST_FEEDER_DRY="1"),
                     stdout=subprocess.PIPE, stderr=subprocess.PIPE).stderr.decode()
m = re.search(r"feeder dry run: ([0-9.]+) Gbp parsed in ([0-9.]+) s with (\d+) threads", dry)
if m:
    res["producers_only"] = {"gbp_per_s": round(float(m.group(1)) / float(m.group(2)), 2), "threads": int(m.group(3))}
res["ours"], of = run(os.path.join(ROOT, "rabbittclust_b200", "host", "bin", "clust-mst"), [], os.path.join(work, "ours"))
res["ours_second_run"], _ = run(exe_ours, [], os.path.join(work, "ours2"))      # CUDA driver files now in the page cache
if not skip_ref:
    nt = os.cpu_count() or 1
    res["reference"], rf = run(os.path.join(ROOT, "oracle", "_ref", "clust-mst"), ["-t", str(nt)], os.path.join(work, "ref"))
    res["reference"]["threads"] = nt
    # with several threads the reference pushes sketches in completion order: compare the sketches as a multiset of rows
    from rabbittclust_b200 import sketch_io
    a = sorted(x.tobytes() for x in sketch_io.read_hash_sketch(os.path.join(rf, "hash.sketch"))["sketches"])
    b = sorted(x.tobytes() for x in sketch_io.read_hash_sketch(os.path.join(of, "hash.sketch"))["sketches"])
    res["same_sketches"] = a == b
print(json.dumps(res))
