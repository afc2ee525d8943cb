"""rabbittclust_b200 -- host-side Python binding of libtclust_cuda.so (include/tclust.h).

The library is the product: hand-written sm_100a CUDA behind a C ABI.  This module only
marshals numpy arrays into it (ctypes) for the tests, bench.py and the Python tools.  There
is no CPU fallback: importing works without a GPU (so the symbol-export tests can run), but
every compute call needs a CUDA device and raises TclustError otherwise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtclust_cuda.so")

EDGE_DTYPE = np.dtype([("pre", "<i4"), ("suf", "<i4"), ("dist", "<f8")])
PAIR_DTYPE = np.dtype([("i", "<u4"), ("j", "<u4"), ("common", "<u4")])

# every symbol include/tclust.h declares (tests check the built library exports all of them)
SYMBOLS = [
    "tc_version", "tc_last_error", "tc_ctx_create", "tc_ctx_destroy", "tc_kernel_launches", "tc_last_pass_kernel_ms", "tc_stream",
    "tc_host_alloc", "tc_host_free", "tc_sketch_minhash", "tc_sketch_minhash_dev", "tc_kssd_shuffled_dim", "tc_sketch_kssd", "tc_sketch_kssd_seq", "tc_sketch_kssd_dev",
    "tc_sketches_upload", "tc_sketches_upload_dev", "tc_sketches_upload_rows", "tc_pair_common", "tc_allpairs_pass",
    "tc_edges_download", "tc_boruvka_upload_comp", "tc_boruvka_phase1", "tc_boruvka_phase2", "tc_boruvka_set_async", "tc_mst",
    "tc_greedy_scan", "tc_greedy_cluster",
    "tc_pair_common_b", "tc_allpairs_pass_b", "tc_mst_bruteforce", "tc_greedy_cluster_bruteforce", "tc_greedy_cluster_kssd", "tc_greedy_incremental_kssd", "tc_greedy_incremental", "tc_dense_ani", "tc_allpairs_pass_from", "tc_mst_append",
    "tc_jaccard_b", "tc_distance_b", "tc_contain_distance",
    "tc_mash_dist", "tc_aaf_dist", "tc_dist_from_key", "tc_radio", "tc_greedy_common_min", "tc_greedy_dist",
    "tc_mg_export", "tc_mg_join", "tc_mg_leave", "tc_mg_rank", "tc_mst_stats", "tc_mst_of_edges", "tc_cluster_threshold", "tc_labels_merge", "tc_greedy_cluster_multi", "tc_kssd_needed_stride", "tc_upload_sort_share_dev", "tc_sketches_upload_runs_dev", "tc_last_h2d_bytes", "tc_vr_partition_dev", "tc_vr_partition32_dev", "tc_vr_local_ranks_dev", "tc_vr_scatter_ranks_dev",
    "tc_sketches_upload_ranks_dev",
]
MG_HANDLE_BYTES = 128


class TclustError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libtclust_cuda.so; fail loudly if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TclustError(f"{LIB_PATH} is missing: build it with `python -m rabbittclust_b200.build` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u8p, u32p, u64p, i32p = (C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_int32))
    L.tc_version.restype = C.c_int
    L.tc_last_error.restype = C.c_char_p
    L.tc_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.tc_ctx_destroy.argtypes = [vp]
    L.tc_ctx_destroy.restype = None
    L.tc_kernel_launches.argtypes = [vp]
    L.tc_kernel_launches.restype = C.c_uint64
    L.tc_last_pass_kernel_ms.argtypes = [vp]
    L.tc_last_pass_kernel_ms.restype = C.c_double
    L.tc_stream.argtypes = [vp]
    L.tc_stream.restype = vp
    L.tc_host_alloc.argtypes = [C.c_size_t]
    L.tc_host_alloc.restype = vp
    L.tc_host_free.argtypes = [vp]
    L.tc_host_free.restype = None
    sk = [vp, vp, u64p, u32p, C.c_uint64, C.c_uint32, C.c_int, u32p, C.c_uint32, vp, C.c_uint64, vp]
    L.tc_sketch_minhash.argtypes = sk
    L.tc_sketch_minhash_dev.argtypes = sk
    L.tc_kssd_shuffled_dim.argtypes = [C.c_int, i32p]
    L.tc_sketch_kssd.argtypes = [vp, vp, u64p, u32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, i32p,
                                 vp, C.c_uint64, vp]
    L.tc_sketch_kssd_dev.argtypes = L.tc_sketch_kssd.argtypes
    L.tc_sketch_kssd_seq.argtypes = L.tc_sketch_kssd.argtypes
    L.tc_sketches_upload.argtypes = [vp, vp, u64p, C.c_uint32]
    L.tc_sketches_upload_dev.argtypes = [vp, vp, u64p, C.c_uint32]
    L.tc_sketches_upload_rows.argtypes = [vp, vp, u64p, C.c_uint32]
    L.tc_pair_common.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u32p]
    L.tc_allpairs_pass.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, u64p, u64p]
    L.tc_edges_download.argtypes = [vp, vp, C.c_uint64, u64p]
    L.tc_boruvka_upload_comp.argtypes = [vp, i32p, C.c_uint32]
    L.tc_boruvka_phase1.argtypes = [vp, C.c_int, vp]
    L.tc_boruvka_phase2.argtypes = [vp, C.c_int, vp, vp]
    L.tc_boruvka_set_async.argtypes = [vp, C.c_int]
    L.tc_mst.argtypes = [vp, C.c_int, C.c_int, C.c_double, vp, u64p]
    L.tc_greedy_scan.argtypes = [vp, C.c_uint32, C.c_uint32, u32p, C.c_uint32, u32p, u32p]
    L.tc_greedy_cluster.argtypes = [vp, C.c_int, C.c_int, C.c_double, i32p, u32p]
    L.tc_pair_common_b.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
    L.tc_allpairs_pass_b.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u64p]
    L.tc_mst_bruteforce.argtypes = [vp, C.c_int, C.c_uint32, C.c_int, vp, u64p]
    L.tc_greedy_cluster_bruteforce.argtypes = [vp, C.c_int, C.c_uint32, C.c_int, C.c_double, i32p, u32p]
    L.tc_greedy_cluster_kssd.argtypes = [vp, C.c_int, C.c_double, i32p, u32p]
    L.tc_greedy_incremental_kssd.argtypes = [vp, C.c_int, C.c_double, C.c_uint32, u32p, C.c_uint32, i32p, u32p]
    L.tc_greedy_incremental.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_uint32, u32p, C.c_uint32, i32p, u32p]
    L.tc_allpairs_pass_from.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u64p]
    L.tc_mst_append.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_uint32, vp, C.c_uint64, vp, u64p]
    L.tc_dense_ani.argtypes = [vp, C.c_int, C.c_int, C.c_int, i32p, u64p]
    L.tc_jaccard_b.argtypes = [C.c_int] * 4
    L.tc_jaccard_b.restype = C.c_double
    L.tc_distance_b.argtypes = [C.c_int] * 5
    L.tc_distance_b.restype = C.c_double
    L.tc_contain_distance.argtypes = [C.c_int] * 4
    L.tc_contain_distance.restype = C.c_double
    for f in (L.tc_mash_dist, L.tc_aaf_dist):
        f.argtypes = [C.c_int] * 4
        f.restype = C.c_double
    L.tc_dist_from_key.argtypes = [C.c_uint64, C.c_int, C.c_int]
    L.tc_dist_from_key.restype = C.c_double
    L.tc_greedy_common_min.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
    L.tc_greedy_common_min.restype = C.c_int
    L.tc_greedy_dist.argtypes = [C.c_int] * 5
    L.tc_greedy_dist.restype = C.c_double
    L.tc_mg_export.argtypes = [vp, C.c_uint32, vp]
    L.tc_mg_join.argtypes = [vp, C.c_int, C.c_int, vp]
    L.tc_mg_leave.argtypes = [vp]
    L.tc_mg_rank.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tc_mst_stats.argtypes = [vp, C.POINTER(C.c_double)]
    L.tc_mst_of_edges.argtypes = [vp, C.c_int, C.c_int, vp, u64p]
    L.tc_cluster_threshold.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_uint32, C.c_uint32, i32p, u32p]
    L.tc_labels_merge.argtypes = [i32p, i32p, C.c_uint32, u32p]
    L.tc_kssd_needed_stride.restype = C.c_uint32
    L.tc_vr_partition_dev.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, u64p, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, u64p]
    L.tc_vr_partition32_dev.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, vp, vp, u64p]
    L.tc_vr_local_ranks_dev.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, u64p]
    L.tc_vr_scatter_ranks_dev.argtypes = [vp, vp, vp, u64p, u64p, C.c_uint32, C.c_uint32, C.c_uint64, vp]
    L.tc_sketches_upload_ranks_dev.argtypes = [vp, vp, u64p, C.c_uint32]
    L.tc_last_h2d_bytes.argtypes = [vp]
    L.tc_last_h2d_bytes.restype = C.c_uint64
    L.tc_upload_sort_share_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, vp]
    L.tc_sketches_upload_runs_dev.argtypes = [vp, vp, vp, u64p, C.c_uint32, u64p, C.c_uint32]
    L.tc_greedy_cluster_multi.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, i32p, u32p]
    L.tc_radio.argtypes = [C.c_double, C.c_int]
    L.tc_radio.restype = C.c_int
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise TclustError(lib().tc_last_error().decode())


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def pack_records(genomes):
    """genomes: list (one per genome) of lists of bytes records -> (seq u8, rec_off u64, rec2genome u32)."""
    recs, r2g = [], []
    for g, rs in enumerate(genomes):
        for r in rs:
            recs.append(r if isinstance(r, (bytes, bytearray)) else r.encode())
            r2g.append(g)
    off = np.zeros(len(recs) + 1, np.uint64)
    if recs:
        off[1:] = np.cumsum([len(r) for r in recs], dtype=np.uint64)
    seq = np.frombuffer(b"".join(recs) + b"\0" * 32, np.uint8).copy()
    return seq, off, np.asarray(r2g, np.uint32)


def pack_sketches(sketches):
    off = np.zeros(len(sketches) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in sketches], dtype=np.uint64)
    parts = [np.asarray(s, np.uint64) for s in sketches] + [np.zeros(1, np.uint64)]
    return np.ascontiguousarray(np.concatenate(parts)), off


class Context:
    """One tc_ctx (one GPU).  Mirrors how the reference's callers use Sketch::MinHash and
    compute_minhash_mst / MinHashGreedyClusterWithInvertedIndex (see include/tclust.h)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _chk(lib().tc_ctx_create(device, C.byref(self._h)))
        self.n = 0
        self.sizes = None
        self._pin, self._pin_bytes = None, 0

    def close(self):
        if self._h:
            if self._pin:
                lib().tc_host_free(self._pin)
                self._pin, self._pin_bytes = None, 0
            lib().tc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().tc_kernel_launches(self._h))

    @property
    def last_pass_kernel_ms(self):
        return float(lib().tc_last_pass_kernel_ms(self._h))

    @property
    def stream(self):
        return int(lib().tc_stream(self._h) or 0)

    # ---- HP-1 ----
    def sketch_minhash(self, genomes, k=21, s=1000, seed=42):
        """genomes: list of lists of records (bytes).  s: int or per-genome sequence.
        Returns list of sorted np.uint64 arrays (Sketch::MinHash::storeMinHashes per genome)."""
        seq, off, r2g = pack_records(genomes)
        ng = len(genomes)
        ss = np.full(ng, s, np.uint32) if np.isscalar(s) else np.asarray(s, np.uint32)
        stride = int(ss.max()) if ng else 1
        out = np.zeros((ng, stride), np.uint64)
        out_n = np.zeros(ng, np.uint32)
        _chk(lib().tc_sketch_minhash(self._h, seq.ctypes.data, _p(off, C.c_uint64), _p(r2g, C.c_uint32),
                                     len(off) - 1, ng, k, _p(ss, C.c_uint32), seed,
                                     out.ctypes.data, stride, out_n.ctypes.data))
        return [out[g, :out_n[g]].copy() for g in range(ng)]

    def sketch_kssd(self, genomes, kmer_size, drlevel, shuffled_dim, max_hashes=None, per_record=False):
        """per_record=True: every genome is one record of a multi-FASTA, sketched by the rule of the single-file
        consumer (tc_sketch_kssd_seq: differs from the list-mode rule for an odd kmer_size)."""
        seq, off, r2g = pack_records(genomes)
        ng = len(genomes)
        if max_hashes is None:
            longest = max((sum(len(r) for r in g) for g in genomes), default=0)
            max_hashes = max(1024, longest // 64)
        out = np.zeros((ng, max_hashes), np.uint64)
        out_n = np.zeros(ng, np.uint32)
        sd = np.ascontiguousarray(shuffled_dim, np.int32)
        fn = lib().tc_sketch_kssd_seq if per_record else lib().tc_sketch_kssd
        _chk(fn(self._h, seq.ctypes.data, _p(off, C.c_uint64), _p(r2g, C.c_uint32),
                len(off) - 1, ng, kmer_size, drlevel, _p(sd, C.c_int32),
                out.ctypes.data, max_hashes, out_n.ctypes.data))
        return [out[g, :out_n[g]].copy() for g in range(ng)]

    # ---- HP-2 ----
    def upload(self, sketches=None, hashes=None, off=None):
        if sketches is not None:
            hashes, off = pack_sketches(sketches)
        hashes = np.ascontiguousarray(hashes, np.uint64)
        off = np.ascontiguousarray(off, np.uint64)
        self.n = len(off) - 1
        self.sizes = np.diff(off).astype(np.int64)
        _chk(lib().tc_sketches_upload(self._h, hashes.ctypes.data, _p(off, C.c_uint64), self.n))

    def upload_rows(self, sketches):
        """The same from one array per sketch, without a flat copy (tc_sketches_upload_rows)."""
        rows = [np.ascontiguousarray(x, np.uint64) for x in sketches]
        self.n = len(rows)
        ptrs = np.array([r.ctypes.data for r in rows], np.uint64)
        sizes = np.array([len(r) for r in rows], np.uint64)
        self.sizes = sizes.astype(np.int64)
        _chk(lib().tc_sketches_upload_rows(self._h, ptrs.ctypes.data, _p(sizes, C.c_uint64), self.n))

    def pair_common(self, i0, i1, j0, j1, impl=0):
        out = np.zeros((i1 - i0, j1 - j0), np.uint32)
        _chk(lib().tc_pair_common(self._h, i0, i1, j0, j1, impl, _p(out, C.c_uint32)))
        return out

    def pair_common_b(self, i0, i1, j0, j1, sketch_size):
        """semantic B (MinHash::jaccard): shared hashes among the first sketch_size union elements."""
        out = np.zeros((i1 - i0, j1 - j0), np.uint32)
        _chk(lib().tc_pair_common_b(self._h, i0, i1, j0, j1, sketch_size, _p(out, C.c_uint32)))
        return out

    def allpairs_pass_b(self, sketch_size, part=0, nparts=1):
        npairs, nedges = C.c_uint64(0), C.c_uint64(0)
        _chk(lib().tc_allpairs_pass_b(self._h, sketch_size, part, nparts, C.byref(npairs), C.byref(nedges)))
        return int(npairs.value), int(nedges.value)

    def mst_bruteforce(self, k, sketch_size, containment=False):
        """modifyMST: MST of the complete graph under MinHash::distance / containDistance."""
        out = np.zeros(max(self.n, 1), EDGE_DTYPE)
        m = C.c_uint64(0)
        _chk(lib().tc_mst_bruteforce(self._h, k, sketch_size, int(containment), out.ctypes.data, C.byref(m)))
        return out[:int(m.value)].copy()

    def greedy_cluster_bruteforce(self, k, sketch_size, threshold, containment=False):
        rep = np.zeros(max(self.n, 1), np.int32)
        nc = C.c_uint32(0)
        _chk(lib().tc_greedy_cluster_bruteforce(self._h, k, sketch_size, int(containment), threshold,
                                                _p(rep, C.c_int32), C.byref(nc)))
        return int(nc.value), rep[:self.n].copy()

    def allpairs_pass_from(self, radio, start_index, part=0, nparts=1):
        npairs, nedges = C.c_uint64(0), C.c_uint64(0)
        _chk(lib().tc_allpairs_pass_from(self._h, radio, start_index, part, nparts, C.byref(npairs), C.byref(nedges)))
        return int(npairs.value), int(nedges.value)

    def mst_append(self, k, threshold, start_index, pre_mst, containment=False):
        """append_clust_mst: MST of (saved MST of the first start_index genomes) + (new genomes vs everything)."""
        pre = np.ascontiguousarray(pre_mst, EDGE_DTYPE)
        out = np.zeros(max(self.n, 1), EDGE_DTYPE)
        m = C.c_uint64(0)
        _chk(lib().tc_mst_append(self._h, k, int(containment), threshold, start_index, pre.ctypes.data, len(pre),
                                 out.ctypes.data, C.byref(m)))
        return out[:int(m.value)].copy()

    def dense_ani(self, k, containment=False, span=100):
        """--dense epilogue over the candidate pairs of the last pass: (dense[span, n] cumulative, ani[101])."""
        dense = np.zeros((span, max(self.n, 1)), np.int32)
        ani = np.zeros(101, np.uint64)
        _chk(lib().tc_dense_ani(self._h, k, int(containment), span, _p(dense, C.c_int32), _p(ani, C.c_uint64)))
        return dense, ani

    def greedy_cluster_kssd(self, k, threshold):
        rep = np.zeros(max(self.n, 1), np.int32)
        nc = C.c_uint32(0)
        _chk(lib().tc_greedy_cluster_kssd(self._h, k, threshold, _p(rep, C.c_int32), C.byref(nc)))
        return int(nc.value), rep[:self.n].copy()

    def greedy_incremental_kssd(self, k, threshold, start_index, rep_ids):
        """KssdIncrementalCluster: sketches [0, start_index) are clustered with representatives rep_ids (creation order);
        returns (#representatives afterwards, rep_of[start_index:n])."""
        rep = np.full(max(self.n, 1), -1, np.int32)
        ids = np.ascontiguousarray(rep_ids, np.uint32)
        nc = C.c_uint32(0)
        _chk(lib().tc_greedy_incremental_kssd(self._h, k, threshold, start_index, _p(ids, C.c_uint32), len(ids), _p(rep, C.c_int32), C.byref(nc)))
        return int(nc.value), rep[start_index:self.n].copy()

    def greedy_incremental(self, k, threshold, start_index, rep_ids, containment=False):
        """MinHashIncrementalCluster: (#representatives afterwards, rep_of[start_index:n])."""
        rep = np.full(max(self.n, 1), -1, np.int32)
        ids = np.ascontiguousarray(rep_ids, np.uint32)
        nc = C.c_uint32(0)
        _chk(lib().tc_greedy_incremental(self._h, k, int(containment), threshold, start_index, _p(ids, C.c_uint32), len(ids),
                                         _p(rep, C.c_int32), C.byref(nc)))
        return int(nc.value), rep[start_index:self.n].copy()

    def allpairs_pass(self, radio, part=0, nparts=1):
        npairs, nedges = C.c_uint64(0), C.c_uint64(0)
        _chk(lib().tc_allpairs_pass(self._h, radio, part, nparts, C.byref(npairs), C.byref(nedges)))
        return int(npairs.value), int(nedges.value)

    def edges(self, pinned=False):
        """Candidate edges of the last pass.  pinned=True: a view of a page-locked buffer the context keeps and
        reuses (valid until the next such call) -- the copy runs at PCIe speed instead of through pageable memory."""
        n = C.c_uint64(0)
        _chk(lib().tc_edges_download(self._h, None, 0, C.byref(n)))
        m = int(n.value)
        if pinned:
            need = max(m, 1) * PAIR_DTYPE.itemsize
            if need > self._pin_bytes:
                if self._pin:
                    lib().tc_host_free(self._pin)
                self._pin_bytes = need + need // 8
                self._pin = lib().tc_host_alloc(self._pin_bytes)
                if not self._pin:
                    self._pin_bytes = 0
                    raise TclustError("tc_host_alloc(%d) failed" % need)
            out = np.frombuffer((C.c_char * (m * PAIR_DTYPE.itemsize)).from_address(self._pin), PAIR_DTYPE, m)
        else:
            out = np.zeros(m, PAIR_DTYPE)
        if m:
            _chk(lib().tc_edges_download(self._h, out.ctypes.data, m, C.byref(n)))
        return out

    def mst(self, k, threshold, containment=False, out=None):
        """out: optional caller-owned array of >= n EDGE_DTYPE entries (the C ABI writes into the caller's buffer); a view
        of its first m entries is returned"""
        if out is None:
            out = np.empty(max(self.n, 1), EDGE_DTYPE)
        assert out.dtype == EDGE_DTYPE and len(out) >= max(self.n, 1) and out.flags.c_contiguous
        m = C.c_uint64(0)
        _chk(lib().tc_mst(self._h, k, int(containment), threshold, out.ctypes.data, C.byref(m)))
        return out[:int(m.value)]

    # ---- multi-GPU group (tc_mg_*): after join(), mst()/mst_append() are collective over the ranks
    def mg_export(self, n_max):
        """Handle (bytes) of this context's exchange buffer, laid out for sets of up to n_max sketches."""
        buf = C.create_string_buffer(MG_HANDLE_BYTES)
        _chk(lib().tc_mg_export(self._h, n_max, buf))
        return buf.raw

    def mg_join(self, rank, nranks, handles):
        """handles: the exported handles of all ranks, in rank order (list of bytes or one bytes object)."""
        blob = handles if isinstance(handles, (bytes, bytearray)) else b"".join(handles)
        assert len(blob) == nranks * MG_HANDLE_BYTES
        _chk(lib().tc_mg_join(self._h, rank, nranks, C.c_char_p(bytes(blob))))

    def mg_leave(self):
        _chk(lib().tc_mg_leave(self._h))

    def mst_of_edges(self, k, mode=0):
        """Boruvka over the candidate edges of the last allpairs_pass (collective in a group)."""
        out = np.zeros(max(self.n, 1), EDGE_DTYPE)
        m = C.c_uint64(0)
        _chk(lib().tc_mst_of_edges(self._h, k, int(mode), out.ctypes.data, C.byref(m)))
        return out[:int(m.value)].copy()

    def mst_stats(self):
        """Timings of the last mst() on this context (CUDA events on its stream), candidate edges, Boruvka rounds."""
        v = (C.c_double * 12)()
        _chk(lib().tc_mst_stats(self._h, v))
        return {"pass_ms": v[0], "boruvka_ms": v[1], "sort_ms": v[2], "edges": int(v[3]), "rounds": int(v[4]),
                "tiles_flagged": int(v[5]), "tiles_total": int(v[6]), "on_reordered_copy": bool(v[7]),
                "bounded_batches": int(v[8]), "bounded_folds": int(v[9])}

    def cluster_threshold(self, k, threshold, containment=False, part=0, nparts=1):
        """Threshold-only clustering (no MST, no edge list): (n_clusters, labels) with labels[v] = smallest id of v's cluster."""
        lab = np.zeros(max(self.n, 1), np.int32)
        nc = C.c_uint32(0)
        _chk(lib().tc_cluster_threshold(self._h, k, int(containment), threshold, part, nparts, _p(lab, C.c_int32), C.byref(nc)))
        return int(nc.value), lab[:self.n]

    def greedy_scan(self, q0, q1, rep_ids):
        reps = np.ascontiguousarray(rep_ids, np.uint32)
        common = np.zeros((q1 - q0, len(reps)), np.uint32)
        first = np.zeros((q1 - q0, len(reps)), np.uint32)
        _chk(lib().tc_greedy_scan(self._h, q0, q1, _p(reps, C.c_uint32), len(reps),
                                  _p(common, C.c_uint32), _p(first, C.c_uint32)))
        return common, first

    def greedy_cluster(self, k, threshold, containment=False):
        rep = np.zeros(max(self.n, 1), np.int32)
        nc = C.c_uint32(0)
        _chk(lib().tc_greedy_cluster(self._h, k, int(containment), threshold, _p(rep, C.c_int32), C.byref(nc)))
        return int(nc.value), rep[:self.n].copy()


def greedy_cluster_multi(ctxs, k, threshold, containment=False, kssd=False):
    """tc_greedy_cluster_multi over contexts that all hold the same uploaded set: (n_clusters, rep_of)."""
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    n = ctxs[0].n
    rep = np.zeros(max(n, 1), np.int32)
    nc = C.c_uint32(0)
    _chk(lib().tc_greedy_cluster_multi(arr, len(ctxs), k, int(containment), threshold, int(kssd), _p(rep, C.c_int32), C.byref(nc)))
    return int(nc.value), rep[:n].copy()


def labels_merge(into, other):
    """Union of two partial partitions given as smallest-id labels (tc_labels_merge); returns (n_clusters, labels)."""
    a = np.ascontiguousarray(into, np.int32).copy()
    b = np.ascontiguousarray(other, np.int32)
    nc = C.c_uint32(0)
    _chk(lib().tc_labels_merge(_p(a, C.c_int32), _p(b, C.c_int32), len(a), C.byref(nc)))
    return int(nc.value), a


def mash_dist(common, size0, size1, k):
    return float(lib().tc_mash_dist(common, size0, size1, k))


def aaf_dist(common, size0, size1, k):
    return float(lib().tc_aaf_dist(common, size0, size1, k))


def distance_b(common_b, size0, size1, sketch_size, k):
    return float(lib().tc_distance_b(common_b, size0, size1, sketch_size, k))


def contain_distance(common, size0, size1, k):
    return float(lib().tc_contain_distance(common, size0, size1, k))


_SHUF = {}


def kssd_shuffled_dim(half_subk=6):
    """generate_shuffle_dim table of the reference (host, glibc rand), cached per half_subk."""
    if half_subk not in _SHUF:
        a = np.zeros(1 << (4 * half_subk), np.int32)
        _chk(lib().tc_kssd_shuffled_dim(half_subk, _p(a, C.c_int32)))
        _SHUF[half_subk] = a
    return _SHUF[half_subk]


def radio(threshold, k):
    return int(lib().tc_radio(threshold, k))
