"""On-disk formats of the reference (the drop-in boundary b-1 of SURVEY.md), Python side.

Layouts follow the reference byte for byte (native little-endian, no padding, no magic):
  hash.sketch  src/Sketch_IO.cpp:192-206 (write), :297-349 (read)
  info.sketch  src/Sketch_IO.cpp:96-131 (write), :477-560 (read)
  edge.mst     src/MST_IO.cpp:200-217 (write), :47-70 (read); info.mst == info.sketch layout
  cluster file src/MST_IO.cpp:136-179
The C++ host (rabbittclust_b200/host/) implements the same formats for the CLI; this module serves
the tests, bench.py and the tools.
"""
import os
import struct

import numpy as np

EDGE_DTYPE = np.dtype([("pre", "<i4"), ("suf", "<i4"), ("dist", "<f8")])


def write_minhash_folder(folder, sketches, names, k, sketch_size=1000, containment=False, contain_compress=1000,
                         total_len=None, comments=None, by_file=True):
    """Write hash.sketch + info.sketch accepted by the reference's --presketched."""
    os.makedirs(folder, exist_ok=True)
    n = len(sketches)
    if total_len is None:
        total_len = [1_000_000] * n
    if comments is None:
        comments = ["synthetic"] * n
    with open(os.path.join(folder, "hash.sketch"), "wb") as fh:
        fh.write(struct.pack("<ii?i", 0, k, bool(containment), contain_compress if containment else sketch_size))
        for h in sketches:
            h = np.ascontiguousarray(h, "<u8")
            fh.write(struct.pack("<Q", len(h)))
            fh.write(h.tobytes())
    with open(os.path.join(folder, "info.sketch"), "wb") as fi:
        fi.write(struct.pack("<?Q", bool(by_file), n))
        for nm, ln, cm in zip(names, total_len, comments):
            sn, cb = nm.encode(), cm.encode()
            if by_file:
                fn = (nm + ".fna").encode()
                fi.write(struct.pack("<iiiiQ", len(fn), len(sn), len(cb), 0, int(ln)))
                fi.write(fn + sn + cb)
            else:
                fi.write(struct.pack("<iiii", len(sn), len(cb), 0, int(ln)))
                fi.write(sn + cb)


def write_minhash_folder_uniform(folder, hashes, n, s, k, total_len=1_000_000):
    """The same two files for n sketches of exactly s hashes each (hashes: flat u64 array of n*s), written in bulk --
    0.8 GB at n = 100000 in a second instead of a Python loop per sketch."""
    os.makedirs(folder, exist_ok=True)
    rec = np.empty((n, s + 1), "<u8")            # rows of [count][hashes]
    rec[:, 0] = s
    rec[:, 1:] = np.asarray(hashes[:n * s], "<u8").reshape(n, s)
    with open(os.path.join(folder, "hash.sketch"), "wb") as fh:
        fh.write(struct.pack("<ii?i", 0, k, False, s))
        rec.tofile(fh)
    with open(os.path.join(folder, "info.sketch"), "wb") as fi:
        fi.write(struct.pack("<?Q", True, n))
        cb = b"synthetic"
        for i in range(n):
            sn = b"g%06d" % i
            fn = sn + b".fna"
            fi.write(struct.pack("<iiiiQ", len(fn), len(sn), len(cb), 0, int(total_len)) + fn + sn + cb)


def read_hash_sketch(path):
    """-> dict(func_id, k, containment, param, sketches=[np.uint64...])."""
    with open(path, "rb") as f:
        data = f.read()
    func_id, k, cont, param = struct.unpack_from("<ii?i", data, 0)
    pos = struct.calcsize("<ii?i")
    sk = []
    while pos < len(data):
        (m,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        sk.append(np.frombuffer(data, "<u8", m, pos).copy())
        pos += 8 * m
    return {"func_id": func_id, "k": k, "containment": bool(cont), "param": param, "sketches": sk}


def read_info_sketch(path):
    """-> (by_file, [dict(file, name, comment, strand, length)])."""
    with open(path, "rb") as f:
        data = f.read()
    by_file, n = struct.unpack_from("<?Q", data, 0)
    pos = 9
    out = []
    for _ in range(n):
        if by_file:
            lf, ls, lc, strand, tot = struct.unpack_from("<iiiiQ", data, pos)
            pos += 24
            fn = data[pos:pos + lf]; pos += lf
            sn = data[pos:pos + ls]; pos += ls
            cm = data[pos:pos + lc]; pos += lc
            out.append({"file": fn.decode(), "name": sn.decode(), "comment": cm.decode(), "strand": strand,
                        "length": tot})
        else:
            ls, lc, strand, length = struct.unpack_from("<iiii", data, pos)
            pos += 16
            sn = data[pos:pos + ls]; pos += ls
            cm = data[pos:pos + lc]; pos += lc
            out.append({"file": "", "name": sn.decode(), "comment": cm.decode(), "strand": strand, "length": length})
    return bool(by_file), out


def read_kssd_hash_sketch(path):
    """kssd.hash.sketch (Sketch_IO.cpp:136-167) -> dict(id, half_k, half_subk, drlevel, genome_number, use64, sketches)."""
    with open(path, "rb") as f:
        data = f.read()
    kid, half_k, half_subk, drlevel, gn = struct.unpack_from("<5i", data, 0)
    use64 = half_k - drlevel > 8
    pos, sk = 20, []
    while pos < len(data):
        (m,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        a = np.frombuffer(data, "<u8" if use64 else "<u4", m, pos)
        pos += (8 if use64 else 4) * m
        sk.append(np.sort(a.astype(np.uint64)))
    return {"id": kid, "half_k": half_k, "half_subk": half_subk, "drlevel": drlevel, "genome_number": gn, "use64": use64,
            "sketches": sk}


def read_kssd_info_sketch(path):
    """kssd.info.sketch (Sketch_IO.cpp:36-86): info.sketch plus one trailing bool use64 per genome -> (by_file, [dict])."""
    with open(path, "rb") as f:
        data = f.read()
    by_file, n = struct.unpack_from("<?Q", data, 0)
    pos, out = 9, []
    for _ in range(n):
        if by_file:
            lf, ls, lc, strand, tot = struct.unpack_from("<iiiiQ", data, pos)
            pos += 24
            fn = data[pos:pos + lf]; pos += lf
            sn = data[pos:pos + ls]; pos += ls
            cm = data[pos:pos + lc]; pos += lc
            out.append({"file": fn.decode(), "name": sn.decode(), "comment": cm.decode(), "strand": strand, "length": tot})
        else:
            ls, lc, strand, length = struct.unpack_from("<iiii", data, pos)
            pos += 16
            sn = data[pos:pos + ls]; pos += ls
            cm = data[pos:pos + lc]; pos += lc
            out.append({"file": "", "name": sn.decode(), "comment": cm.decode(), "strand": strand, "length": length})
        pos += 1
    return bool(by_file), out


def read_kssd_cluster_state(path):
    """cluster_state.bin of clust-greedy --fast --save-rep (KssdClusterState::save, greedy.cpp:1545-1625) ->
    dict(threshold, kmer_size, half_k, half_subk, drlevel, genome_number, rep_ids, sketches=[dict(id, length, use64,
    hashes, file)], clusters, index={hash: [representative numbers]})."""
    with open(path, "rb") as f:
        data = f.read()
    thr, k, half_k, half_subk, drlevel, gn = struct.unpack_from("<d5i", data, 0)
    pos = 28
    (nrep,) = struct.unpack_from("<Q", data, pos); pos += 8
    rep_ids = list(struct.unpack_from("<%di" % nrep, data, pos)); pos += 4 * nrep
    (nsk,) = struct.unpack_from("<Q", data, pos); pos += 8
    sk = []
    for _ in range(nsk):
        gid, length, use64, ssize, n32, n64 = struct.unpack_from("<iQ?IQQ", data, pos)
        pos += struct.calcsize("<iQ?IQQ")
        h32 = np.frombuffer(data, "<u4", n32, pos); pos += 4 * n32
        h64 = np.frombuffer(data, "<u8", n64, pos); pos += 8 * n64
        (nl,) = struct.unpack_from("<Q", data, pos); pos += 8
        fn = data[pos:pos + nl].decode(); pos += nl
        sk.append({"id": gid, "length": length, "use64": bool(use64), "file": fn,
                   "hashes": np.sort((h64 if n64 else h32).astype(np.uint64))})
    (ncl,) = struct.unpack_from("<Q", data, pos); pos += 8
    clusters = []
    for _ in range(ncl):
        (m,) = struct.unpack_from("<Q", data, pos); pos += 8
        clusters.append(list(struct.unpack_from("<%di" % m, data, pos))); pos += 4 * m
    index = {}
    wide = data[pos:pos + 6] == b"KSSI02"
    if wide:
        pos += 8
    (nidx,) = struct.unpack_from("<Q", data, pos); pos += 8
    for _ in range(nidx):
        if wide:
            (h,) = struct.unpack_from("<Q", data, pos); pos += 8
        else:
            (h,) = struct.unpack_from("<I", data, pos); pos += 4
        (m,) = struct.unpack_from("<Q", data, pos); pos += 8
        index[h] = list(struct.unpack_from("<%di" % m, data, pos)); pos += 4 * m
    assert pos == len(data), "trailing bytes in " + path
    return {"threshold": thr, "kmer_size": k, "half_k": half_k, "half_subk": half_subk, "drlevel": drlevel,
            "genome_number": gn, "rep_ids": rep_ids, "sketches": sk, "clusters": clusters, "index": index}


def read_minhash_cluster_state(path):
    """cluster_state.bin of clust-greedy --save-rep (MinHashClusterState::save, greedy.cpp:2134-2205) ->
    dict(threshold, kmer_size, sketch_size, containment, rep_ids, sketches=[dict(id, length, hashes, file)], clusters,
    index={hash: [representative numbers]})."""
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b"MINHASH\0", "not a MinHash cluster state: " + path
    thr, k, ssz, cont = struct.unpack_from("<dii?", data, 8)
    pos = 8 + 17
    (nrep,) = struct.unpack_from("<Q", data, pos); pos += 8
    rep_ids = list(struct.unpack_from("<%di" % nrep, data, pos)); pos += 4 * nrep
    (nsk,) = struct.unpack_from("<Q", data, pos); pos += 8
    sk = []
    for _ in range(nsk):
        gid, length, nh = struct.unpack_from("<iQQ", data, pos); pos += 20
        h = np.frombuffer(data, "<u8", nh, pos).copy(); pos += 8 * nh
        (nl,) = struct.unpack_from("<Q", data, pos); pos += 8
        fn = data[pos:pos + nl].decode(); pos += nl
        sk.append({"id": gid, "length": length, "hashes": h, "file": fn})
    (ncl,) = struct.unpack_from("<Q", data, pos); pos += 8
    clusters = []
    for _ in range(ncl):
        (m,) = struct.unpack_from("<Q", data, pos); pos += 8
        clusters.append(list(struct.unpack_from("<%di" % m, data, pos))); pos += 4 * m
    (nidx,) = struct.unpack_from("<Q", data, pos); pos += 8
    index = {}
    for _ in range(nidx):
        h, m = struct.unpack_from("<QQ", data, pos); pos += 16
        index[h] = list(struct.unpack_from("<%di" % m, data, pos)); pos += 4 * m
    assert pos == len(data), "trailing bytes in " + path
    return {"threshold": thr, "kmer_size": k, "sketch_size": ssz, "containment": bool(cont), "rep_ids": rep_ids,
            "sketches": sk, "clusters": clusters, "index": index}


def read_edge_mst(path):
    with open(path, "rb") as f:
        data = f.read()
    (n,) = struct.unpack_from("<Q", data, 0)
    return np.frombuffer(data, EDGE_DTYPE, n, 8).copy()


def write_edge_mst(path, edges):
    e = np.ascontiguousarray(edges, EDGE_DTYPE)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(e)))
        f.write(e.tobytes())


def parse_cluster_file(path):
    """Cluster text (MST_IO.cpp:136-179) -> list of clusters, each a list of (global_id, name) tuples."""
    clusters, cur = [], None
    for line in open(path):
        if line.startswith("#"):
            continue
        if line.startswith("the cluster"):
            cur = []
            clusters.append(cur)
        elif line.startswith("\t") and cur is not None:
            f = line.rstrip("\n").split("\t")
            # list mode: '', idx, gid, len, file, name, comment ; single-file mode: '', idx, gid, len, name, comment
            gid = int(f[2])
            name = f[5].strip() if len(f) >= 7 else f[4].strip()
            cur.append((gid, name))
    return clusters
