#!/usr/bin/env python
"""Golden vectors of the single-file KSSD consumer (consumer_fasta_task_with_kssd, src/SketchInfo.cpp:274-433) from the
UNMODIFIED reference (oracle/_ref/clust-mst --fast -i): one multi-FASTA with N runs of many lengths, lower case, other
ambiguity codes, valid runs of exactly k-2 .. k+2 bases and records shorter than k, sketched per record
  * with ODD k-mer sizes (13 = what tune_kssd_parameters recommends for these lengths, and 15: the window is used after k
    valid bases while the tuple holds 2*half_k = k+1 bases, so the first window after every reset carries one stale base,
    and the hash is packed with a shift derived from the odd k),
  * with an EVEN one (16: the windows of the list-mode loop),
all at --drlevel 0 (every window passes the dimension filter, so every window is visible in the sketch).
Writes tests/golden/kssd_seq/{multi.fa, k13/kssd.hash.sketch, k13/kssd.info.sketch, k15/..., k16/...}.
usage: python tools/make_golden_kssd_seq.py   (needs oracle/_ref; rewrites tests/golden/kssd_seq)"""
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden", "kssd_seq")


def records():
    rng = np.random.default_rng(99)
    acgt = np.frombuffer(b"ACGT", np.uint8)

    def rnd(n):
        return bytes(acgt[rng.integers(0, 4, n)])
    recs = []
    s = bytearray(rnd(2500))                 # scattered N runs of various lengths, a lower-case stretch
    p = 50
    while p < 2400:
        L = int(rng.choice([1, 1, 2, 3, 7, 20, 120]))
        s[p:p + L] = b"N" * L
        p += L + int(rng.integers(5, 60))
    s[1000:1100] = bytes(s[1000:1100]).lower()
    recs.append(("r0 scattered N runs", bytes(s)))
    s = bytearray(rnd(1500))                 # starts and ends with N, other ambiguity codes
    s[0:5] = b"NNNNN"; s[-3:] = b"NNN"; s[500] = ord("R"); s[520] = ord("y"); s[521] = ord("K"); s[900:919] = b"N" * 19
    recs.append(("r1", bytes(s)))
    parts = []                               # valid runs of k-2 .. k+2 bases for k = 13, 15, 16 between Ns
    for L in (11, 12, 13, 14, 15, 16, 17, 18, 14, 15, 16, 40, 13, 15):
        parts.append(rnd(L)); parts.append(b"N" * int(rng.integers(1, 4)))
    recs.append(("r2 short runs", b"".join(parts) + rnd(800)))
    recs.append(("r3 plain", rnd(1500)))
    for L in (12, 13, 14, 15, 16, 17, 5):    # whole records around the k-mer sizes
        recs.append(("short%d" % L, rnd(L)))
    return recs


def main():
    if os.path.exists(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    fa = os.path.join(GOLD, "multi.fa")
    with open(fa, "wb") as f:
        for n, s in records():
            f.write(b">" + n.encode() + b"\n")
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + b"\n")
    meta = {}
    for k in (13, 15, 16):
        with tempfile.TemporaryDirectory() as tmp:
            cmd = [os.path.join(REF, "clust-mst"), "-t", "1", "--fast", "-i", fa, "-k", str(k), "--drlevel", "0", "-m", "1",
                   "-d", "0.05", "-o", "out.txt"]
            p = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            if p.returncode:
                sys.exit("FAILED: %s\n%s" % (" ".join(cmd), p.stderr.decode()[-3000:]))
            folder = [d for d in glob.glob(os.path.join(tmp, "20*")) if os.path.isdir(d)][0]
            out = os.path.join(GOLD, "k%d" % k)
            os.makedirs(out)
            for fn in ("kssd.hash.sketch", "kssd.info.sketch"):
                shutil.copy(os.path.join(folder, fn), out)
            err = p.stderr.decode()
            meta["k%d" % k] = {"cmd": " ".join(["oracle/_ref/clust-mst"] + cmd[1:]).replace(fa, "tests/golden/kssd_seq/multi.fa"),
                               "stderr_kmer_lines": [l for l in err.splitlines() if "kmer" in l.lower() or "half_k" in l][:8]}
    json.dump(meta, open(os.path.join(GOLD, "MANIFEST.json"), "w"), indent=1)
    subprocess.call(["du", "-sh", GOLD])
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
