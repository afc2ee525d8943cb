#!/usr/bin/env python
"""Golden vectors of the reference's incremental greedy rules on sets full of EXACT distance ties
(rabbittclust_b200.synth.incremental_tie_set): the winner among equal distances is the first candidate in the iteration
order of the reference's `phmap::flat_hash_map<int,int> candidate_counts` -- reproduced by the oracle and the product.
  tests/golden/incremental/kssd.json        KssdInitialClusterWithState + KssdIncrementalCluster   (ref_harness kssdinc)
  tests/golden/incremental/minhash_s.json   MinHash..., Mash mode                                   (ref_harness minhashinc)
  tests/golden/incremental/minhash_c.json   MinHash..., containment mode
  tests/golden/incremental/phmap_order.json insertion sequences -> iteration order of the real container (ref_harness phorder)
  tests/golden/incremental/greedy_kssd_full.json   the whole-set rule KssdGreedyClusterWithInvertedIndex: the unmodified
                                            clust-greedy --fast --presketched (-t 1) on synth.kssd_like_families sets
                                            (many equal sizes and Jaccard values): processing order + clusters
Each file holds the generator arguments (the sketches are regenerated from the seed), the state order, the representative
ids and the clusters the UNMODIFIED reference produced with one thread.
usage: python tools/make_golden_incremental.py   (needs oracle/_ref; rewrites tests/golden/incremental)"""
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabbittclust_b200 import sketch_io, synth  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
GOLD = os.path.join(ROOT, "tests", "golden", "incremental")


def write_kssd_folder(folder, sketches, half_k=9, half_subk=6, drlevel=3):
    os.makedirs(folder, exist_ok=True)
    n = len(sketches)
    with open(os.path.join(folder, "kssd.hash.sketch"), "wb") as f:
        f.write(struct.pack("<iiiii", (half_k << 8) + (half_subk << 4) + drlevel, half_k, half_subk, drlevel, n))
        for h in sketches:
            f.write(struct.pack("<Q", len(h)))
            f.write(np.asarray(h).astype("<u4").tobytes())
    with open(os.path.join(folder, "kssd.info.sketch"), "wb") as f:
        f.write(struct.pack("<?Q", True, n))
        for i in range(n):
            fn, nm, cm = b"f%05d.fna" % i, b"s%05d" % i, b"c"
            f.write(struct.pack("<iiiiQ", len(fn), len(nm), len(cm), 0, 2000000))
            f.write(fn + nm + cm)
            f.write(struct.pack("<?", False))


def run_harness(args):
    p = subprocess.run([HARNESS] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if p.returncode:
        sys.exit("FAILED: ref_harness %s\n%s" % (" ".join(args), p.stderr.decode()[-2000:]))
    lines = p.stdout.decode().splitlines()
    order = [int(x) for x in lines[0].split()]
    reps = [int(x) for x in lines[1].split()]
    clusters = [[] if l.strip() == "-" else [int(x) for x in l.split()] for l in lines[2:]]
    return order, reps, clusters


def main():
    if os.path.exists(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    cases = {
        "kssd": dict(gen=dict(seed=1, n_reps=2000, n_queries=400, size=99, threshold=0.08, kmer=18), cmd="kssdinc"),
        "minhash_s": dict(gen=dict(seed=2, n_reps=1900, n_queries=400, size=99, threshold=0.08, kmer=21, hash_bits=62), cmd="minhashinc"),
        "minhash_c": dict(gen=dict(seed=3, n_reps=600, n_queries=300, size=99, query_size=60, threshold=0.08, kmer=21, containment=True,
                                   hash_bits=62), cmd="minhashinc"),
    }
    for name, c in cases.items():
        sk, n_old = synth.incremental_tie_set(**c["gen"])
        with tempfile.TemporaryDirectory() as tmp:
            if c["cmd"] == "kssdinc":
                write_kssd_folder(tmp, sk)
            else:
                contain = c["gen"].get("containment", False)
                sketch_io.write_minhash_folder(tmp, sk, ["g%05d" % i for i in range(len(sk))], c["gen"]["kmer"],
                                               sketch_size=c["gen"]["size"], containment=contain, contain_compress=1000)
            order, reps, clusters = run_harness([c["cmd"], tmp, str(n_old), str(c["gen"]["threshold"])])
        n_new_reps = len(reps) - sum(1 for r in reps if r < n_old)
        multi = sum(1 for cl in clusters if len(cl) > 1)
        json.dump({"generator": c["gen"], "harness": "ref_harness %s <folder> %d %s" % (c["cmd"], n_old, c["gen"]["threshold"]),
                   "n_old": n_old, "order": order, "rep_ids": reps, "clusters": clusters},
                  open(os.path.join(GOLD, name + ".json"), "w"))
        print(name, "sketches", len(sk), "representatives", len(reps), "(new: %d)" % n_new_reps, "clusters with members", multi)
    # ---- KssdGreedyClusterWithInvertedIndex through the unmodified binary
    full = []
    for gen in (dict(seed=3, n_fam=12, per_fam=10), dict(seed=4, n_fam=30, per_fam=12, keep_lo=0.3), dict(seed=5, n_fam=6, per_fam=40, pool=200, keep_lo=0.4)):
        sk = synth.kssd_like_families(**gen)
        with tempfile.TemporaryDirectory() as tmp:
            folder = os.path.join(tmp, "kssd")
            write_kssd_folder(folder, sk)
            p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "clust-greedy"), "-t", "1", "--fast", "--presketched", folder, "-d", "0.05",
                                "-o", "out.txt"], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            if p.returncode:
                sys.exit("FAILED clust-greedy --fast --presketched\n" + p.stderr.decode()[-2000:])
            ref = sketch_io.parse_cluster_file(os.path.join(tmp, "out.txt"))
        # the cluster file carries (position in the processing order, name s<folder index>): the std::sort by size is not stable
        pos_of = {}
        for c in ref:
            for gid, nm in c:
                pos_of[gid] = int(nm[1:])
        assert sorted(pos_of) == list(range(len(sk)))
        full.append({"generator": gen, "cmd": "oracle/_ref/clust-greedy -t 1 --fast --presketched <folder> -d 0.05 -o out.txt", "kmer": 18,
                     "threshold": 0.05, "order": [pos_of[p] for p in range(len(sk))], "clusters": [[gid for gid, _ in c] for c in ref]})
        print("greedy_kssd_full", gen, "clusters", len(ref))
    json.dump(full, open(os.path.join(GOLD, "greedy_kssd_full.json"), "w"))
    # the candidate table itself
    rng = np.random.default_rng(5)
    seqs = []
    for n, hi in ((5, 100), (300, 1000), (1791, 100000), (1792, 100000), (1793, 100000), (2500, 3000), (3700, 1000000)):
        keys = rng.choice(hi, n, replace=False).astype(np.int32)
        with tempfile.TemporaryDirectory() as tmp:
            keys.tofile(os.path.join(tmp, "k.bin"))
            subprocess.check_call([HARNESS, "phorder", os.path.join(tmp, "k.bin"), os.path.join(tmp, "o.bin")])
            seqs.append({"keys": keys.tolist(), "order": np.fromfile(os.path.join(tmp, "o.bin"), np.int32).tolist()})
    json.dump(seqs, open(os.path.join(GOLD, "phmap_order.json"), "w"))
    subprocess.call(["du", "-sh", GOLD])


if __name__ == "__main__":
    main()
