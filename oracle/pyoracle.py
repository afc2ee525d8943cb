"""ctypes binding of oracle/liboracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from rabbittclust_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Edge(C.Structure):
    _fields_ = [("pre", C.c_int32), ("suf", C.c_int32), ("dist", C.c_double)]


EDGE_DTYPE = np.dtype([("pre", "<i4"), ("suf", "<i4"), ("dist", "<f8")])


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _DIR, "all"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _DIR, "oracle"])
        L = C.CDLL(path)
        u8p, u64p, i32p, u32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_uint32))
        L.oc_murmur3_x64_128_h1.restype = C.c_uint64
        L.oc_murmur3_x64_128_h1.argtypes = [u8p, C.c_int, C.c_uint32]
        L.oc_minhash_sketch.restype = C.c_size_t
        L.oc_minhash_sketch.argtypes = [u8p, u64p, C.c_int, C.c_int, C.c_size_t, C.c_uint32, C.c_int, u64p]
        L.oc_kssd_shuffled_dim.restype = None
        L.oc_kssd_shuffled_dim.argtypes = [C.c_int, i32p]
        L.oc_kssd_sketch.restype = C.c_size_t
        L.oc_kssd_sketch.argtypes = [u8p, u64p, C.c_int, C.c_int, C.c_int, i32p, u64p, C.c_size_t]
        L.oc_kssd_sketch_seq.restype = C.c_size_t
        L.oc_kssd_sketch_seq.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, i32p, u64p, C.c_size_t]
        L.oc_intersect_u64.restype = C.c_uint32
        L.oc_intersect_u64.argtypes = [u64p, C.c_size_t, u64p, C.c_size_t, u32p]
        L.oc_jaccard_B.restype = C.c_double
        L.oc_jaccard_B.argtypes = [u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint64]
        L.oc_distance_B.restype = C.c_double
        L.oc_distance_B.argtypes = [C.c_double, C.c_int]
        L.oc_radio.restype = C.c_int
        L.oc_radio.argtypes = [C.c_double, C.c_int]
        for f in (L.oc_mash_dist_A, L.oc_aaf_dist_mst):
            f.restype = C.c_double
            f.argtypes = [C.c_int] * 4
        L.oc_greedy_dist.restype = C.c_double
        L.oc_greedy_dist.argtypes = [C.c_int] * 5
        L.oc_greedy_common_min.restype = C.c_int
        L.oc_greedy_common_min.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oc_mst.restype = C.c_size_t
        L.oc_mst.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(Edge)]
        L.oc_common_matrix.restype = None
        L.oc_common_matrix.argtypes = [u64p, u64p, C.c_int, u32p]
        L.oc_clusters_from_mst.restype = C.c_int
        L.oc_clusters_from_mst.argtypes = [C.POINTER(Edge), C.c_size_t, C.c_int, C.c_double, i32p, i32p]
        L.oc_greedy_inverted.restype = C.c_int
        L.oc_greedy_inverted.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_int, C.c_double, i32p]
        L.oc_greedy_incremental_kssd.restype = C.c_int
        L.oc_greedy_incremental_kssd.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_double, C.c_int, i32p, C.c_int, i32p]
        L.oc_greedy_kssd.restype = C.c_int
        L.oc_greedy_kssd.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_double, i32p]
        L.oc_phmap_iteration_order.restype = None
        L.oc_phmap_iteration_order.argtypes = [i32p, C.c_int, i32p]
        L.oc_greedy_incremental_minhash.restype = C.c_int
        L.oc_greedy_incremental_minhash.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, i32p, C.c_int, i32p]
        L.oc_common_B.restype = C.c_uint32
        L.oc_common_B.argtypes = [u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint64, u32p]
        L.oc_contain_distance.restype = C.c_double
        L.oc_contain_distance.argtypes = [u64p, C.c_size_t, u64p, C.c_size_t, C.c_int]
        L.oc_mst_bruteforce.restype = C.c_size_t
        L.oc_mst_bruteforce.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(Edge)]
        L.oc_greedy_bruteforce.restype = C.c_int
        L.oc_greedy_bruteforce.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_double, i32p]
        L.oc_dense_ani.restype = None
        L.oc_dense_ani.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, i32p, u64p]
        L.oc_remove_noise_clusters.restype = C.c_int
        L.oc_remove_noise_clusters.argtypes = [C.POINTER(Edge), C.c_size_t, C.c_int, C.c_double, i32p, C.c_int, i32p, i32p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def murmur_h1(key: bytes, seed=42):
    a = np.frombuffer(key, np.uint8)
    return int(lib().oc_murmur3_x64_128_h1(_p(a, C.c_uint8), len(key), seed))


def _pack_records(records):
    recs = [r if isinstance(r, (bytes, bytearray)) else r.encode() for r in records]
    off = np.zeros(len(recs) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in recs])
    seq = np.frombuffer(b"".join(recs) + b"\0", np.uint8).copy()
    return seq, off, len(recs)


def minhash_sketch(records, k=21, s=1000, seed=42, rule=0):
    """records: list of bytes/str (FASTA records of ONE genome) -> sorted np.uint64."""
    seq, off, n = _pack_records(records)
    out = np.zeros(max(int(s), 1), np.uint64)
    m = lib().oc_minhash_sketch(_p(seq, C.c_uint8), _p(off, C.c_uint64), n, k, s, seed, rule,
                                _p(out, C.c_uint64))
    return out[:m].copy()


_SHUF = {}


def kssd_shuffled_dim(half_subk=6):
    if half_subk not in _SHUF:
        a = np.zeros(1 << (4 * half_subk), np.int32)
        lib().oc_kssd_shuffled_dim(half_subk, _p(a, C.c_int32))
        _SHUF[half_subk] = a
    return _SHUF[half_subk]


def kssd_params(kmer_size, drlevel=3):
    half_k = (kmer_size + 1) // 2
    half_subk = 6 if 6 - drlevel >= 2 else drlevel + 2
    return half_k, half_subk, drlevel


def kssd_sketch(records, kmer_size, drlevel=3):
    seq, off, n = _pack_records(records)
    _, half_subk, _ = kssd_params(kmer_size, drlevel)
    sd = kssd_shuffled_dim(half_subk)
    cap = max(1024, int(off[-1]) // (8 if drlevel else 1) + 1024)      # drlevel 0 keeps every window
    out = np.zeros(cap, np.uint64)
    m = lib().oc_kssd_sketch(_p(seq, C.c_uint8), _p(off, C.c_uint64), n, kmer_size, drlevel,
                             _p(sd, C.c_int32), _p(out, C.c_uint64), cap)
    return out[:m].copy()


def kssd_sketch_seq(record, kmer_size, drlevel=3):
    """one record of a multi-FASTA through the single-file consumer (odd k: stale-base windows, un-rounded shift)"""
    seq = np.frombuffer(bytes(record), np.uint8).copy() if len(record) else np.zeros(1, np.uint8)
    _, half_subk, _ = kssd_params(kmer_size, drlevel)
    sd = kssd_shuffled_dim(half_subk)
    cap = len(record) + 16
    out = np.zeros(cap, np.uint64)
    m = lib().oc_kssd_sketch_seq(_p(seq, C.c_uint8), len(record), kmer_size, drlevel, _p(sd, C.c_int32), _p(out, C.c_uint64), cap)
    return out[:m].copy()


def pack_sketches(sketches):
    off = np.zeros(len(sketches) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in sketches])
    h = np.concatenate([np.asarray(s, np.uint64) for s in sketches] + [np.zeros(1, np.uint64)])
    return np.ascontiguousarray(h), off


def intersect(a, b):
    a = np.ascontiguousarray(a, np.uint64); b = np.ascontiguousarray(b, np.uint64)
    fh = C.c_uint32(0)
    c = lib().oc_intersect_u64(_p(a, C.c_uint64), len(a), _p(b, C.c_uint64), len(b), C.byref(fh))
    return int(c), int(fh.value)


def jaccard_B(a, b, sketch_size):
    a = np.ascontiguousarray(a, np.uint64); b = np.ascontiguousarray(b, np.uint64)
    return float(lib().oc_jaccard_B(_p(a, C.c_uint64), len(a), _p(b, C.c_uint64), len(b), sketch_size))


def common_B(a, b, sketch_size):
    """-> (common, denom) of MinHash::jaccard (semantic B)."""
    a = np.ascontiguousarray(a, np.uint64); b = np.ascontiguousarray(b, np.uint64)
    d = C.c_uint32(0)
    c = lib().oc_common_B(_p(a, C.c_uint64), len(a), _p(b, C.c_uint64), len(b), sketch_size, C.byref(d))
    return int(c), int(d.value)


def common_B_matrix(sketches, sketch_size):
    n = len(sketches)
    out = np.zeros((n, n), np.uint32)
    for i in range(n):
        for j in range(n):
            out[i, j] = common_B(sketches[i], sketches[j], sketch_size)[0]
    return out


def distance_B(a, b, sketch_size, k):
    return float(lib().oc_distance_B(jaccard_B(a, b, sketch_size), k))


def contain_distance(a, b, k):
    a = np.ascontiguousarray(a, np.uint64); b = np.ascontiguousarray(b, np.uint64)
    return float(lib().oc_contain_distance(_p(a, C.c_uint64), len(a), _p(b, C.c_uint64), len(b), k))


def mst_bruteforce(sketches, k, sketch_size, containment=False):
    h, off = pack_sketches(sketches)
    n = len(sketches)
    out = np.zeros(max(n, 1), EDGE_DTYPE)
    m = lib().oc_mst_bruteforce(_p(h, C.c_uint64), _p(off, C.c_uint64), n, k, sketch_size, int(containment),
                                out.ctypes.data_as(C.POINTER(Edge)))
    return out[:m].copy()


def greedy_bruteforce(sketches, k, sketch_size, threshold, containment=False):
    h, off = pack_sketches(sketches)
    n = len(sketches)
    rep = np.zeros(max(n, 1), np.int32)
    nc = lib().oc_greedy_bruteforce(_p(h, C.c_uint64), _p(off, C.c_uint64), n, k, sketch_size, int(containment),
                                    threshold, _p(rep, C.c_int32))
    return nc, rep[:n]


def dense_ani(sketches, k, threshold, containment=False, span=100):
    h, off = pack_sketches(sketches)
    n = len(sketches)
    dense = np.zeros((span, max(n, 1)), np.int32)
    ani = np.zeros(101, np.uint64)
    lib().oc_dense_ani(_p(h, C.c_uint64), _p(off, C.c_uint64), n, k, int(containment), threshold, span,
                       _p(dense, C.c_int32), _p(ani, C.c_uint64))
    return dense, ani


def remove_noise_clusters(edges, n, threshold, dense):
    e = np.ascontiguousarray(edges, EDGE_DTYPE)
    d = np.ascontiguousarray(dense, np.int32)
    labels = np.zeros(max(n, 1), np.int32); order = np.zeros(max(n, 1), np.int32)
    nc = lib().oc_remove_noise_clusters(e.ctypes.data_as(C.POINTER(Edge)), len(e), n, threshold, _p(d, C.c_int32),
                                        d.shape[0], _p(labels, C.c_int32), _p(order, C.c_int32))
    return nc, labels[:n], order[:n]


def common_matrix(sketches):
    h, off = pack_sketches(sketches)
    n = len(sketches)
    out = np.zeros((n, n), np.uint32)
    lib().oc_common_matrix(_p(h, C.c_uint64), _p(off, C.c_uint64), n, _p(out, C.c_uint32))
    return out


def mst(sketches, k, threshold, containment=False):
    h, off = pack_sketches(sketches)
    n = len(sketches)
    out = np.zeros(max(n, 1), EDGE_DTYPE)
    m = lib().oc_mst(_p(h, C.c_uint64), _p(off, C.c_uint64), n, k, int(containment), threshold,
                     out.ctypes.data_as(C.POINTER(Edge)))
    return out[:m].copy()


def clusters_from_mst(edges, n, threshold):
    e = np.ascontiguousarray(edges, EDGE_DTYPE)
    labels = np.zeros(max(n, 1), np.int32); order = np.zeros(max(n, 1), np.int32)
    nc = lib().oc_clusters_from_mst(e.ctypes.data_as(C.POINTER(Edge)), len(e), n, threshold,
                                    _p(labels, C.c_int32), _p(order, C.c_int32))
    return nc, labels[:n], order[:n]


def greedy(sketches, k, threshold, containment=False):
    h, off = pack_sketches(sketches)
    n = len(sketches)
    rep = np.zeros(max(n, 1), np.int32)
    nc = lib().oc_greedy_inverted(_p(h, C.c_uint64), _p(off, C.c_uint64), n, k, int(containment),
                                  threshold, _p(rep, C.c_int32))
    return nc, rep[:n]


def greedy_incremental_kssd(sketches, k, threshold, start, seeds):
    """KssdIncrementalCluster: (number of representatives afterwards, rep_of[start:])."""
    h, off = pack_sketches(sketches)
    n = len(sketches)
    rep = np.full(max(n, 1), -1, np.int32)
    sd = np.ascontiguousarray(seeds, np.int32)
    nc = lib().oc_greedy_incremental_kssd(_p(h, C.c_uint64), _p(off, C.c_uint64), n, k, threshold, start,
                                          _p(sd, C.c_int32), len(sd), _p(rep, C.c_int32))
    return nc, rep[start:n]


def greedy_kssd(sketches, k, threshold):
    """KssdGreedyClusterWithInvertedIndex on sketches already in the reference's processing order: (#clusters, rep_of)"""
    h, off = pack_sketches(sketches)
    n = len(sketches)
    rep = np.zeros(max(n, 1), np.int32)
    nc = lib().oc_greedy_kssd(_p(h, C.c_uint64), _p(off, C.c_uint64), n, k, threshold, _p(rep, C.c_int32))
    return nc, rep[:n]


def phmap_iteration_order(keys):
    """iteration order of the reference's candidate table for distinct int keys given in insertion order"""
    k = np.ascontiguousarray(keys, np.int32)
    out = np.zeros(max(len(k), 1), np.int32)
    lib().oc_phmap_iteration_order(_p(k, C.c_int32), len(k), _p(out, C.c_int32))
    return out[:len(k)]


def greedy_incremental_minhash(sketches, k, threshold, start, seeds, containment=False):
    """MinHashIncrementalCluster: (number of representatives afterwards, rep_of[start:])."""
    h, off = pack_sketches(sketches)
    n = len(sketches)
    rep = np.full(max(n, 1), -1, np.int32)
    sd = np.ascontiguousarray(seeds, np.int32)
    nc = lib().oc_greedy_incremental_minhash(_p(h, C.c_uint64), _p(off, C.c_uint64), n, k, int(containment), threshold, start,
                                             _p(sd, C.c_int32), len(sd), _p(rep, C.c_int32))
    return nc, rep[start:n]


def ref_bin(name):
    """Path of a prebuilt reference binary under oracle/_ref, or None."""
    p = os.path.join(_DIR, "_ref", name)
    return p if os.path.exists(p) else None
