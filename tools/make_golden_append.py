#!/usr/bin/env python
"""Golden vectors of the classic --append flows, produced by the UNMODIFIED reference (oracle/_ref) in this container:
the first 9 genomes of tests/golden/mst_genomes are clustered, the other 4 appended to that folder.
  tests/golden/append/first.txt, second.txt          the two genome lists (paths relative to the repo root)
  tests/golden/append/kssd/base, .../result          clust-mst --fast, then --fast --append   (sub_command.cpp:1286-1477)
  tests/golden/append/minhash/base, .../result       clust-mst -k 15 -s 1000, then --append   (sub_command.cpp:1578-1757)
  tests/golden/append/greedy_kssd/base, step1, step2 clust-greedy --fast on the first 9, then --fast --append --save-rep of
                                                     second_a.txt (4 genomes; builds cluster_state.bin) and of second_b.txt
                                                     (3 genomes; continues from the state)   (sub_command.cpp:192-276)
  tests/golden/append/greedy_minhash/{s,c}           clust-greedy -k 15 (-s 1000 | -c 100) --save-rep on the first 9, then
                                                     --append --save-rep of second.txt from the saved state
                                                     (sub_command.cpp:97-190, MinHashIncrementalCluster); the states are kept as
                                                     JSON summaries (the sketches are those of tests/golden/append/minhash)
usage: python tools/make_golden_append.py   (needs oracle/_ref; rewrites tests/golden/append)"""
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden", "append")
N_FIRST = 9


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if p.returncode:
        sys.exit("FAILED: %s\n%s" % (" ".join(cmd), p.stderr.decode()[-2000:]))
    return p


def newest_dir(cwd, known):
    d = [p for p in glob.glob(os.path.join(cwd, "20*")) if os.path.isdir(p) and p not in known]
    assert len(d) == 1, d
    return d[0]


def main():
    if os.path.exists(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    files = open(os.path.join(ROOT, "tests", "golden", "mst_genomes", "list.txt")).read().split()
    open(os.path.join(GOLD, "first.txt"), "w").write("\n".join(files[:N_FIRST]) + "\n")
    open(os.path.join(GOLD, "second.txt"), "w").write("\n".join(files[N_FIRST:]) + "\n")
    meta = {}
    flows = {
        "kssd": (["--fast"], ["kssd.hash.sketch", "kssd.info.sketch", "kssd.info.mst", "edge.mst", "kssd.sketch.index", "kssd.sketch.dict"]),
        "minhash": (["-k", "15", "-s", "1000"], ["hash.sketch", "info.sketch", "info.mst", "edge.mst"]),
    }
    for name, (flags, keep) in flows.items():
        with tempfile.TemporaryDirectory() as tmp:
            os.symlink(os.path.join(ROOT, "tests"), os.path.join(tmp, "tests"))
            exe = os.path.join(REF, "clust-mst")
            base_cmd = [exe, "-t", "1"] + flags + ["-l", "-i", "tests/golden/append/first.txt", "-d", "0.05", "-o", "base.txt"]
            run(base_cmd, tmp)
            base = newest_dir(tmp, [])
            # upstream bug: the MinHash append path reads minhash.sketch.index with an incompatible loader
            idx = os.path.join(base, "minhash.sketch.index")
            if os.path.exists(idx):
                os.remove(idx)
            out_base = os.path.join(GOLD, name, "base")
            os.makedirs(out_base)
            for fn in keep:
                shutil.copy(os.path.join(base, fn), out_base)
            shutil.copy(os.path.join(tmp, "base.txt"), out_base)
            fast = ["--fast"] if name == "kssd" else []
            app_cmd = [exe, "-t", "1"] + fast + ["-l", "--append", "tests/golden/append/second.txt", "--presketched", base,
                                                  "-d", "0.05", "-o", "app.txt"]
            p = run(app_cmd, tmp)
            res = newest_dir(tmp, [base])
            out_res = os.path.join(GOLD, name, "result")
            os.makedirs(out_res)
            for fn in keep:
                if not fn.startswith("kssd.sketch."):
                    shutil.copy(os.path.join(res, fn), out_res)
            shutil.copy(os.path.join(tmp, "app.txt"), out_res)
            meta[name] = {"base_cmd": " ".join(["oracle/_ref/clust-mst"] + base_cmd[1:]),
                          "append_cmd": " ".join(["oracle/_ref/clust-mst"] + app_cmd[1:]).replace(base, "<base folder>"),
                          "stderr_tail": p.stderr.decode()[-400:]}
    # ---- clust-greedy --fast --append with the cluster state (append_clust_greedy_fast, KssdIncrementalCluster)
    open(os.path.join(GOLD, "second_a.txt"), "w").write("\n".join(files[N_FIRST:N_FIRST + 4]) + "\n")
    open(os.path.join(GOLD, "second_b.txt"), "w").write("\n".join(files[N_FIRST + 4:]) + "\n")
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(ROOT, "tests"), os.path.join(tmp, "tests"))
        exe = os.path.join(REF, "clust-greedy")
        base_cmd = [exe, "-t", "1", "--fast", "-l", "-i", "tests/golden/append/first.txt", "-d", "0.05", "-o", "base.txt"]
        run(base_cmd, tmp)
        base = newest_dir(tmp, [])
        out = os.path.join(GOLD, "greedy_kssd")
        os.makedirs(os.path.join(out, "base"))
        for fn in ("kssd.hash.sketch", "kssd.info.sketch"):
            shutil.copy(os.path.join(base, fn), os.path.join(out, "base"))
        shutil.copy(os.path.join(tmp, "base.txt"), os.path.join(out, "base"))
        known = [base]
        cmds = []
        for step, lst in (("step1", "second_a.txt"), ("step2", "second_b.txt")):
            cmd = [exe, "-t", "1", "--fast", "-l", "--append", "tests/golden/append/" + lst, "--presketched", base, "--save-rep",
                   "-d", "0.05", "-o", step + ".txt"]
            p = run(cmd, tmp)
            new = newest_dir(tmp, known)            # the sketches of the appended genomes (compute_kssd_sketches saves them)
            known.append(new)
            os.makedirs(os.path.join(out, step, "new"))
            for fn in ("kssd.hash.sketch", "kssd.info.sketch"):
                shutil.copy(os.path.join(new, fn), os.path.join(out, step, "new"))
            shutil.copy(os.path.join(base, "cluster_state.bin"), os.path.join(out, step))
            shutil.copy(os.path.join(tmp, step + ".txt"), os.path.join(out, step, "app.txt"))
            cmds.append(" ".join(["oracle/_ref/clust-greedy"] + cmd[1:]).replace(base, "<base folder>"))
        meta["greedy_kssd"] = {"base_cmd": " ".join(["oracle/_ref/clust-greedy"] + base_cmd[1:]), "append_cmds": cmds,
                               "stderr_tail": p.stderr.decode()[-400:]}
    # ---- clust-greedy --append from a saved MinHash cluster state (append_clust_greedy, MinHashIncrementalCluster)
    sys.path.insert(0, ROOT)
    from rabbittclust_b200 import sketch_io

    def summary(path):
        st = sketch_io.read_minhash_cluster_state(path)
        want = {}
        for r, gid in enumerate(st["rep_ids"]):
            for h in st["sketches"][gid]["hashes"]:
                want.setdefault(int(h), []).append(r)
        assert want == st["index"]                  # the index section is what the representatives' sketches imply
        return {"threshold": st["threshold"], "kmer_size": st["kmer_size"], "sketch_size": st["sketch_size"],
                "containment": st["containment"], "rep_ids": st["rep_ids"], "clusters": st["clusters"],
                "sketches": [[x["id"], x["length"], len(x["hashes"]), x["file"]] for x in st["sketches"]]}
    for tag, flags in (("s", ["-s", "1000"]), ("c", ["-c", "100"])):
        with tempfile.TemporaryDirectory() as tmp:
            os.symlink(os.path.join(ROOT, "tests"), os.path.join(tmp, "tests"))
            exe = os.path.join(REF, "clust-greedy")
            base_cmd = [exe, "-t", "1", "-l", "-i", "tests/golden/append/first.txt", "-k", "15"] + flags + ["-d", "0.05", "--save-rep", "-o", "base.txt"]
            run(base_cmd, tmp)
            base = newest_dir(tmp, [])
            out = os.path.join(GOLD, "greedy_minhash", tag)
            os.makedirs(out)
            for fn in ("hash.sketch", "info.sketch"):
                shutil.copy(os.path.join(base, fn), out)
            shutil.copy(os.path.join(tmp, "base.txt"), out)
            json.dump(summary(os.path.join(base, "cluster_state.bin")), open(os.path.join(out, "state_base.json"), "w"))
            app_cmd = [exe, "-t", "1", "-l", "--append", "tests/golden/append/second.txt", "--presketched", base, "--save-rep",
                       "-d", "0.05", "-o", "app.txt"]
            p = run(app_cmd, tmp)
            shutil.copy(os.path.join(tmp, "app.txt"), out)
            json.dump(summary(os.path.join(base, "cluster_state.bin")), open(os.path.join(out, "state_app.json"), "w"))
            meta["greedy_minhash_" + tag] = {"base_cmd": " ".join(["oracle/_ref/clust-greedy"] + base_cmd[1:]),
                                             "append_cmd": " ".join(["oracle/_ref/clust-greedy"] + app_cmd[1:]).replace(base, "<base folder>"),
                                             "stderr_tail": p.stderr.decode()[-300:]}
    json.dump(meta, open(os.path.join(GOLD, "MANIFEST.json"), "w"), indent=1)
    subprocess.call(["du", "-sh", GOLD])


if __name__ == "__main__":
    main()
