#!/usr/bin/env python
"""Golden vectors of the --dense path from the UNMODIFIED reference (oracle/_ref/clust-mst --dense -t 1):
tests/golden/dense/{mash,contain}/ = hash.sketch, info.sketch, edge.mst, mst.ani, mst.dense, clusters.txt,
clusters.txt.removeNoise of a list-mode run on tests/golden/mst_genomes, and presk_{mash,contain}.removeNoise = the
noise-removed cluster files of --presketched runs on tests/golden/presketched (the containment one removes 3 genomes);
tests/golden/trees/ = cluster file, Newick/PHYLIP/NEXUS trees and linkage matrix of `--premsted tests/golden/mst_list`.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "clust-mst")
DST = os.path.join(ROOT, "tests", "golden", "dense")


def main():
    shutil.rmtree(DST, ignore_errors=True)
    for tag, extra in (("mash", ["-s", "1000"]), ("contain", ["-c", "100"])):
        out = os.path.join(DST, tag)
        os.makedirs(out)
        with tempfile.TemporaryDirectory() as tmp:
            os.symlink(os.path.join(ROOT, "tests"), os.path.join(tmp, "tests"))
            subprocess.run([REF, "--dense", "-t", "1", "-l", "-i", "tests/golden/mst_genomes/list.txt", "-k", "15", "-d", "0.05",
                            "-o", "clusters.txt"] + extra, cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            d = [p for p in glob.glob(os.path.join(tmp, "20*")) if os.path.isdir(p)][0]
            for fn in ("hash.sketch", "info.sketch", "edge.mst", "mst.ani", "mst.dense"):
                shutil.copy(os.path.join(d, fn), out)
            for fn in ("clusters.txt", "clusters.txt.removeNoise"):
                shutil.copy(os.path.join(tmp, fn), out)
        with tempfile.TemporaryDirectory() as tmp:
            shutil.copytree(os.path.join(ROOT, "tests", "golden", "presketched", tag), os.path.join(tmp, "p"))
            subprocess.run([REF, "--dense", "-t", "1", "--presketched", "p", "-d", "0.05", "-o", "c.txt"], cwd=tmp, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            shutil.copy(os.path.join(tmp, "c.txt.removeNoise"), os.path.join(DST, "presk_%s.removeNoise" % tag))
    # tree / linkage outputs of --premsted on the committed MST folder (tests/golden/trees)
    trees = os.path.join(ROOT, "tests", "golden", "trees")
    shutil.rmtree(trees, ignore_errors=True)
    os.makedirs(trees)
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([REF, "-t", "1", "--premsted", os.path.join(ROOT, "tests", "golden", "mst_list"), "-d", "0.05", "-o", "mst.txt",
                        "--newick-tree", "--phylip-tree", "--nexus-tree", "--linkage-matrix"], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for suffix in ("", ".newick.tree", ".phylip.tree", ".nexus.tree", ".linkage.txt"):
            shutil.copy(os.path.join(tmp, "mst.txt" + suffix), trees)
    subprocess.call(["du", "-sh", DST, trees])


if __name__ == "__main__":
    main()
