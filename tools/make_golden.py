#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference binaries in oracle/_ref (built by
oracle/Makefile from /root/reference).  Run in the build container only; the fixtures are
committed so that the tests on the GPU box never need the reference.

Fixtures:
  genomes/*.fa.gz            small synthetic genomes: plain, lower-case, N runs, IUPAC, multi-record,
                             a record shorter than k
  ecoli/*.fa.gz              first 150 kb of the reference's own bundled E. coli fixtures
                             (RabbitSketch/examples/genome{1,2,3}.fna) -- data, not source
  sketches.npz               ref_harness (Sketch::MinHash) sketches of all of the above, k=21 s=1000,
                             plus k=15 s=200 and k=32 s=64
  mst_list/                  clust-mst -l run (-t 1) on mst_genomes/*.fa.gz: hash.sketch, info.sketch,
                             edge.mst, cluster text
  greedy_list/               clust-greedy -l run (-t 1), Mash mode (-s) and containment (-c)
  presketched/               synthetic sketch folder written by sketch_io + the reference's cluster text
                             for clust-mst and clust-greedy --presketched
  kssd/                      clust-mst --fast -l run: kssd.hash.sketch + cluster text
  semb/                      semantic-B vectors (tools/make_golden_semb.py): MinHash::jaccard/distance/
                             containDistance of all pairs, modifyMST weights, greedyCluster centres
"""
import glob
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabbittclust_b200 import sketch_io, synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")
ECOLI = "/root/reference/RabbitSketch/examples"


def write_fasta(path, records, width=80):
    with gzip.open(path, "wb", compresslevel=9) as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + b"\n")


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if p.returncode:
        sys.stderr.write(p.stderr.decode()[-3000:])
        raise SystemExit("failed: " + " ".join(cmd))
    return p


def harness_sketch(k, s, files):
    out = subprocess.check_output([os.path.join(REF, "ref_harness"), "sketch", str(k), str(s)] + files)
    a = np.frombuffer(out, np.uint64)
    res, pos = [], 0
    for _ in files:
        n = int(a[pos]); res.append(a[pos + 1:pos + 1 + n].copy()); pos += 1 + n
    return res


def newest_dir(cwd):
    ds = [d for d in glob.glob(os.path.join(cwd, "20*")) if os.path.isdir(d)]
    return max(ds, key=os.path.getmtime)


def main():
    rng = np.random.default_rng(2024)
    shutil.rmtree(GOLD, ignore_errors=True)
    os.makedirs(os.path.join(GOLD, "genomes"))
    os.makedirs(os.path.join(GOLD, "ecoli"))

    # ---- 1. sketch fixtures
    g_plain = synth.random_genome(60000, rng)
    g_mut = synth.mutate(g_plain, 0.01, rng)
    g_lower = bytes(g_plain[:30000]).lower() + g_plain[30000:]
    gn = bytearray(synth.random_genome(50000, rng))
    for s0 in (1000, 20000, 20021, 33333):
        gn[s0:s0 + int(rng.integers(1, 60))] = b"N" * 60
    for p in rng.integers(0, 50000, 80):
        gn[p] = b"RYKMSWnry"[int(rng.integers(0, 9))]
    g_multi = [("r%d" % i, synth.random_genome(int(L), rng)) for i, L in enumerate([5000, 12, 20, 21, 22, 8000, 1, 3000])]
    files = {}
    for name, recs in [("plain", [("plain", g_plain)]), ("mut", [("mut", g_mut)]), ("lower", [("lower", g_lower)]),
                       ("nrun", [("nrun", bytes(gn))]), ("multi", g_multi)]:
        p = os.path.join(GOLD, "genomes", name + ".fa.gz")
        write_fasta(p, recs)
        files[name] = p
    for i in (1, 2, 3):
        seq = b"".join(l.strip() for l in open(os.path.join(ECOLI, "genome%d.fna" % i), "rb") if not l.startswith(b">"))
        p = os.path.join(GOLD, "ecoli", "ecoli%d_150k.fa.gz" % i)
        write_fasta(p, [("ecoli%d first 150 kb" % i, seq[:150000])])
        files["ecoli%d" % i] = p
    names = sorted(files)
    npz = {}
    for (k, s) in [(21, 1000), (15, 200), (32, 64), (17, 1000)]:
        sk = harness_sketch(k, s, [files[n] for n in names])
        for n, h in zip(names, sk):
            npz["%s_k%d_s%d" % (n, k, s)] = h
    np.savez_compressed(os.path.join(GOLD, "sketches.npz"), **npz)

    # ---- 2. clust-mst / clust-greedy from genomes (list mode, -t 1)
    gen_dir = os.path.join(GOLD, "mst_genomes")
    os.makedirs(gen_dir)
    fam = synth.genome_families(4, 4, 30000, seed=99, rates=(0.0, 0.002, 0.01, 0.03))
    rel = []
    for i, g in enumerate(fam):
        nm = "g%02d" % i
        write_fasta(os.path.join(gen_dir, nm + ".fa.gz"), [(nm + " family %d member %d" % (i // 4, i % 4), g)])
        rel.append("tests/golden/mst_genomes/%s.fa.gz" % nm)
    with open(os.path.join(gen_dir, "list.txt"), "w") as f:
        f.write("\n".join(rel) + "\n")
    meta = {}
    for tag, exe, extra in [("mst_list", "clust-mst", ["-s", "1000"]),
                            ("greedy_list_s", "clust-greedy", ["-s", "1000"]),
                            ("greedy_list_c", "clust-greedy", ["-c", "100"])]:
        out = os.path.join(GOLD, tag)
        os.makedirs(out)
        with tempfile.TemporaryDirectory() as tmp:
            os.symlink(os.path.join(ROOT, "tests"), os.path.join(tmp, "tests"))
            cmd = [os.path.join(REF, exe), "-t", "1", "-l", "-i", "tests/golden/mst_genomes/list.txt", "-k", "15",
                   "-d", "0.05", "-o", "clusters.txt"] + extra
            p = run(cmd, tmp)
            d = newest_dir(tmp)
            for fn in os.listdir(d):
                if fn in ("hash.sketch", "info.sketch", "edge.mst", "info.mst"):
                    shutil.copy(os.path.join(d, fn), out)
            shutil.copy(os.path.join(tmp, "clusters.txt"), out)
            meta[tag] = {"cmd": " ".join(["oracle/_ref/" + exe] + cmd[1:]),
                         "stderr_tail": p.stderr.decode()[-400:]}

    # ---- 3. presketched folders
    pres = os.path.join(GOLD, "presketched")
    sk = synth.family_sketches(6, 7, s=300, seed=5)
    sketch_io.write_minhash_folder(os.path.join(pres, "mash"), sk, ["s%03d" % i for i in range(len(sk))], k=21,
                                   sketch_size=300, total_len=[100000 + 13 * i for i in range(len(sk))])
    skc = synth.family_sketches(5, 6, seed=6, ragged=(100, 900))
    sketch_io.write_minhash_folder(os.path.join(pres, "contain"), skc, ["c%03d" % i for i in range(len(skc))], k=21,
                                   containment=True, contain_compress=1000,
                                   total_len=[1000 * len(s) + i for i, s in enumerate(skc)])
    for mode in ("mash", "contain"):
        for exe in ("clust-mst", "clust-greedy"):
            with tempfile.TemporaryDirectory() as tmp:
                cmd = [os.path.join(REF, exe), "-t", "1", "--presketched", os.path.join(pres, mode), "-d", "0.05",
                       "-o", "clusters.txt"]
                run(cmd, tmp)
                shutil.copy(os.path.join(tmp, "clusters.txt"), os.path.join(pres, mode, exe + ".clusters.txt"))
        for fn in os.listdir(os.path.join(pres, mode)):   # drop index caches the reference wrote into the folder
            if fn.endswith(".index") or fn.endswith(".dict"):
                os.remove(os.path.join(pres, mode, fn))

    # ---- 4. KSSD (--fast) list mode
    out = os.path.join(GOLD, "kssd")
    os.makedirs(out)
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(ROOT, "tests"), os.path.join(tmp, "tests"))
        cmd = [os.path.join(REF, "clust-mst"), "--fast", "-t", "1", "-l", "-i", "tests/golden/mst_genomes/list.txt",
               "-d", "0.05", "-o", "clusters.txt"]
        p = run(cmd, tmp)
        d = newest_dir(tmp)
        for fn in os.listdir(d):
            if fn in ("kssd.hash.sketch", "kssd.info.sketch", "edge.mst"):
                shutil.copy(os.path.join(d, fn), out)
        shutil.copy(os.path.join(tmp, "clusters.txt"), out)
        meta["kssd"] = {"cmd": " ".join(cmd[1:]), "stderr_tail": p.stderr.decode()[-600:]}
    # ---- 5. semantic-B brute-force paths (MinHash::distance, modifyMST, greedyCluster)
    import make_golden_semb
    make_golden_semb.main()
    meta["semb"] = {"cmd": "oracle/_ref/ref_harness pairsdump|mstb|greedyb (tools/make_golden_semb.py)"}
    json.dump(meta, open(os.path.join(GOLD, "MANIFEST.json"), "w"), indent=1)
    subprocess.call(["du", "-sh", GOLD])


if __name__ == "__main__":
    main()
