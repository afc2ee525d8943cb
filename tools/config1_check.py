#!/usr/bin/env python
"""BASELINE.json configs[0] at full size: clust-mst on 1,000 synthetic 1 Mbp genomes (50 families x 20, substitution
rates cycled over 0.001..0.05), k=21 s=1000 -d 0.05, list mode.  Runs the UNMODIFIED reference
(oracle/_ref/clust-mst -t 1) and our binary on the same files and compares: hash.sketch / info.sketch byte-identical,
sorted MST weights bit-identical, same partition.  Prints one JSON line with the phase times of both.
With --ref-threads T the reference also runs with -t T (its sketch order then depends on the thread schedule, so only
its wall/phase times are recorded from that run, not compared).
usage: python tools/config1_check.py [workdir] [n_families] [per_family] [length] [--ref-threads T]"""
import glob
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


def run(exe, args, cwd):
    os.makedirs(cwd, exist_ok=True)
    t0 = time.time()
    r = subprocess.run([exe] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    dt = time.time() - t0
    if r.returncode:
        sys.stderr.write(r.stderr.decode()[-2000:])
        raise SystemExit("failed: %s" % exe)
    folder = [p for p in glob.glob(os.path.join(cwd, "20*")) if os.path.isdir(p)][0]
    err = r.stderr.decode()
    phases = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"=+time of ([^=]+?) is: ([0-9.eE+-]+)", err)}
    return folder, dt, phases, err


def main():
    ref_threads = 0
    if "--ref-threads" in sys.argv:
        i = sys.argv.index("--ref-threads"); ref_threads = int(sys.argv[i + 1]); del sys.argv[i:i + 2]
    work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/config1"
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    per = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    length = int(sys.argv[4]) if len(sys.argv) > 4 else 1000000
    gdir = os.path.join(work, "genomes")
    os.makedirs(gdir, exist_ok=True)
    t0 = time.time()
    gen = synth.genome_families(nf, per, length, seed=12345)
    files = []
    for i, g in enumerate(gen):
        f, m = divmod(i, per)
        p = os.path.join(gdir, "g%04d_%03d.fna" % (f, m))
        a = np.frombuffer(g, np.uint8)
        rows = a[: len(a) // 80 * 80].reshape(-1, 80)
        body = np.concatenate([rows, np.full((len(rows), 1), 10, np.uint8)], axis=1).tobytes() + bytes(a[len(a) // 80 * 80:]) + b"\n"
        with open(p, "wb") as fh:
            fh.write((">g%04d_%03d synthetic family %d member %d\n" % (f, m, f, m)).encode() + body)
        files.append(p)
    lst = os.path.join(work, "list.txt")
    open(lst, "w").write("\n".join(files) + "\n")
    gen_s = time.time() - t0
    args = ["-l", "-i", lst, "-k", "21", "-s", "1000", "-d", "0.05", "-o", "out.txt"]
    rf, rt, rph, rerr = run(os.path.join(ROOT, "oracle", "_ref", "clust-mst"), ["-t", "1"] + args, os.path.join(work, "ref"))
    of, ot, oph, oerr = run(os.path.join(ROOT, "rabbittclust_b200", "host", "bin", "clust-mst"), args, os.path.join(work, "ours"))
    same = lambda fn: open(os.path.join(rf, fn), "rb").read() == open(os.path.join(of, fn), "rb").read()
    a = sketch_io.read_edge_mst(os.path.join(rf, "edge.mst"))
    b = sketch_io.read_edge_mst(os.path.join(of, "edge.mst"))
    part = lambda p: sorted(sorted(n for _, n in c) for c in sketch_io.parse_cluster_file(p))
    res = {"genomes": len(files), "length": length, "generate_s": gen_s,
           "k_replaced": ("replace the kmerSize" in rerr, "replace the kmerSize" in oerr),
           "hash_sketch_identical": same("hash.sketch"), "info_sketch_identical": same("info.sketch"),
           "mst_weights_identical": bool(np.array_equal(np.sort(a["dist"]), np.sort(b["dist"]))),
           "partition_identical": part(os.path.join(work, "ref", "out.txt")) == part(os.path.join(work, "ours", "out.txt")),
           "clusters": len(part(os.path.join(work, "ours", "out.txt"))),
           "reference_t1": {"wall_s": rt, "phases_s": rph}, "ours": {"wall_s": ot, "phases_s": oph}}
    res["ours"]["timing_lines"] = [l for l in oerr.splitlines() if l.startswith("[tclust timing]")]
    if ref_threads > 1:
        _, rt2, rph2, _ = run(os.path.join(ROOT, "oracle", "_ref", "clust-mst"), ["-t", str(ref_threads)] + args, os.path.join(work, "ref_mt"))
        res["reference_mt"] = {"threads": ref_threads, "wall_s": rt2, "phases_s": rph2}
        _, ot2, oph2, _ = run(os.path.join(ROOT, "rabbittclust_b200", "host", "bin", "clust-mst"), args, os.path.join(work, "ours2"))
        res["ours_second_run"] = {"wall_s": ot2, "phases_s": oph2}
    print(json.dumps(res))
    assert res["hash_sketch_identical"] and res["info_sketch_identical"] and res["mst_weights_identical"] and res["partition_identical"]


if __name__ == "__main__":
    main()
