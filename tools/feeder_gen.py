#!/usr/bin/env python
"""Writes G synthetic genomes of L bp (one 80-column FASTA file each) under <workdir>/genomes and <workdir>/list.txt
(input of tools/feeder_check.py and of the feeder dry run: TCLUST_FEEDER_DRY=1 clust-mst -l -i list.txt).
usage: python tools/feeder_gen.py <workdir> <genomes> <length>"""
import os, sys
import numpy as np


def generate(work, G, L, seed=3):
    gdir = os.path.join(work, "genomes")
    os.makedirs(gdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    files = []
    for g in range(G):
        a = acgt[rng.integers(0, 4, L)]
        rows = a[: L // 80 * 80].reshape(-1, 80)
        body = np.concatenate([rows, np.full((len(rows), 1), 10, np.uint8)], axis=1).tobytes() + bytes(a[L // 80 * 80:]) + b"\n"
        p = os.path.join(gdir, "g%05d.fna" % g)
        with open(p, "wb") as fh:
            fh.write((">g%05d synthetic\n" % g).encode() + body)
        files.append(p)
    lst = os.path.join(work, "list.txt")
    with open(lst, "w") as fh:
        fh.write("\n".join(files) + "\n")
    return lst


if __name__ == "__main__":
    print(generate(sys.argv[1], int(sys.argv[2]), int(sys.argv[3])))
