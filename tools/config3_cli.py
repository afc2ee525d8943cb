#!/usr/bin/env python
"""BASELINE.json configs[2] through the drop-in binary: a --presketched folder of N synthetic s=1000 sketches
(families of 100, the bench model) -> clust-mst --presketched -d 0.05.  Checks that the clusters are the planted
families and prints one JSON line with wall and phase times; with --ref also runs the unmodified reference binary
(oracle/_ref/clust-mst -t THREADS) on the same folder and compares the partitions.
usage: python tools/config3_cli.py [--n 100000] [--work DIR] [--ref] [--threads T] [--skip-ours]"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabbittclust_b200 import sketch_io, synth  # noqa: E402


def write_folder(folder, h, off, k=21, s=1000):
    os.makedirs(folder, exist_ok=True)
    n = len(off) - 1
    rec = np.empty((n, s + 1), "<u8")           # every sketch has exactly s hashes: [len][hashes] rows
    assert np.all(np.diff(off) == s)
    rec[:, 0] = s
    rec[:, 1:] = h.reshape(n, s)
    with open(os.path.join(folder, "hash.sketch"), "wb") as fh:
        import struct
        fh.write(struct.pack("<ii?i", 0, k, False, s))
        fh.write(rec.tobytes())
    names = ["g%06d" % i for i in range(n)]
    import struct
    with open(os.path.join(folder, "info.sketch"), "wb") as fi:
        fi.write(struct.pack("<?Q", True, n))
        for nm in names:
            fn, sn, cb = (nm + ".fna").encode(), nm.encode(), b"synthetic"
            fi.write(struct.pack("<iiiiQ", len(fn), len(sn), len(cb), 0, 1000000))
            fi.write(fn + sn + cb)


def run(exe, args, cwd):
    os.makedirs(cwd, exist_ok=True)
    t0 = time.time()
    r = subprocess.run([exe] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    dt = time.time() - t0
    err = r.stderr.decode()
    if r.returncode:
        sys.stderr.write(err[-3000:])
        raise SystemExit("failed: %s" % exe)
    sys.stderr.write("".join(l + "\n" for l in err.splitlines() if l.startswith("[tclust timing]")))
    phases = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"=+time of ([^=]+?) is: ([0-9.eE+-]+)", err)}
    return dt, phases


def partition(path):
    clusters = sketch_io.parse_cluster_file(path)
    return sorted(tuple(sorted(m[0] if isinstance(m, tuple) else m for m in c)) for c in clusters)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--work", default="/tmp/config3")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--skip-ours", action="store_true")
    a = ap.parse_args()
    folder = os.path.join(a.work, "sketches")
    t0 = time.time()
    h, off = synth.family_sketches_fast(a.n // 100, 100, 1000, seed=7)
    write_folder(folder, h, off)
    out = {"n": a.n, "generate_s": round(time.time() - t0, 2)}
    args = ["--presketched", folder, "-d", "0.05", "-o", "out.txt"]
    parts = {}
    if not a.skip_ours:
        dt, ph = run(os.path.join(ROOT, "rabbittclust_b200", "host", "bin", "clust-mst"), args, os.path.join(a.work, "ours"))
        out["ours"] = {"wall_s": round(dt, 3), "phases": ph}
        parts["ours"] = partition(os.path.join(a.work, "ours", "out.txt"))
        out["ours"]["clusters"] = len(parts["ours"])
        out["ours"]["planted_families_recovered"] = parts["ours"] == sorted(tuple(range(f * 100, f * 100 + 100)) for f in range(a.n // 100))
    if a.ref:
        dt, ph = run(os.path.join(ROOT, "oracle", "_ref", "clust-mst"), ["-t", str(a.threads)] + args, os.path.join(a.work, "ref"))
        out["reference"] = {"wall_s": round(dt, 3), "threads": a.threads, "phases": ph}
        parts["ref"] = partition(os.path.join(a.work, "ref", "out.txt"))
        out["reference"]["clusters"] = len(parts["ref"])
        if "ours" in parts:
            out["same_partition"] = parts["ours"] == parts["ref"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
