"""Multi-GPU MST build: the tile space of the dense pass is sharded over the ranks of one box and the
only exchange is the Boruvka min-edge reduction (two all-reduce(MIN) of N u64 words per round).

Replaces compute_minhash_mst / compute_kssd_mst (reference src/MST.cpp:1290-1737, 216-807) for
world_size > 1.  One process per GPU (torchrun); torch.distributed is plumbing only: the dense pass and
both reduction phases are the CUDA kernels behind include/tclust.h.  Every rank ends with the same MST.

The per-rank compute is injected (`ops`) so that the host-side logic here -- tile ownership, the
two-phase min reduction, component relabelling, distance conversion -- is testable on CPU with the gloo
backend (tests/test_dist_cpu.py supplies an oracle-backed `ops`).
"""
import ctypes as C

import numpy as np

U64_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)
_SIGN = np.int64(-2**63)


def tile_owner(t, nparts):
    """Rank that owns tile t of the row-major lower-triangular 32x32-group tile order (the rule of
    k3_tile_kernel: the k-th tile a rank fetches is k*nparts + part)."""
    return t % nparts


def join_cseg(n, nparts=1, num_sms=148, start_index=0):
    """Columns per unit of the default dense pass (k3h_join.cu: kj_pick_cseg): the widest power of two up to 2^20
    (a unit is then a whole row of tiles) that leaves every CTA of every rank 2 units; 4096 at least (and always with
    TCLUST_PASS_IMPL=join)."""
    nrb, rb0 = (n + 31) // 32, start_index // 32
    cseg = 1 << 20
    while cseg > 4096:
        units = sum(nrb - max(c0 // 32, rb0) for c0 in range(0, n, cseg))
        if units >= 2 * nparts * num_sms:
            return cseg
        cseg >>= 1
    return 4096


def join_units(n, cseg=None, nparts=1):
    """Work units of the default dense pass (k3h_join.cu: tc_launch_blockjoin_edges): row block rb x column
    segment [c0, c1), segment-major.  Unit u belongs to rank u % nparts."""
    if cseg is None:
        cseg = join_cseg(n, nparts)
    nrb = (n + 31) // 32
    units = []
    for c0 in range(0, n, cseg):
        for rb in range(c0 // 32, nrb):
            c1 = min(n, c0 + cseg, rb * 32 + 31)
            if c1 > c0:
                units.append((rb, c0, c1))
    return units


def owned_pairs(n, part, nparts, impl="join", group=32):
    """Number of pairs j<i owned by `part` (mirror of tc_allpairs_pass's n_pairs) for the hash-join pass
    (units) or the first-generation merge pass (32x32 tiles, TCLUST_PASS_IMPL=merge)."""
    total = 0
    if impl == "join":
        for u, (rb, c0, c1) in enumerate(join_units(n, nparts=nparts)):
            if u % nparts != part:
                continue
            for i in range(rb * 32, min(n, rb * 32 + 32)):
                total += max(0, min(c1, i) - c0)
        return total
    g = (n + group - 1) // group
    t = 0
    for r in range(g):
        nr = min(group, n - r * group)
        for c in range(r + 1):
            if t % nparts == part:
                nc = min(group, n - c * group)
                total += nr * (nr - 1) // 2 if r == c else nr * nc
            t += 1
    return total


def allreduce_min_u64(arr, dist=None, device=None):
    """Element-wise unsigned-64 minimum over all ranks.  NCCL/gloo reduce signed int64, so the sign bit
    is flipped around the reduction (order-preserving map u64 -> i64)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return arr
    import torch
    t = torch.from_numpy((arr.view(np.int64) ^ _SIGN).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return (t.cpu().numpy() ^ _SIGN).view(np.uint64)


class GpuOps:
    """Boruvka reduction phases of one rank on its GPU (tc_boruvka_* over the rank's candidate edges)."""

    def __init__(self, ctx, containment):
        import torch
        import rabbittclust_b200 as T
        self.T, self.torch, self.ctx, self.containment = T, torch, ctx, int(containment)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.key = torch.empty(ctx.n, dtype=torch.int64, device=self.dev)
        self.edge = torch.empty(ctx.n, dtype=torch.int64, device=self.dev)

    def phase1(self, comp):
        L, T = self.T.lib(), self.T
        T._chk(L.tc_boruvka_upload_comp(self.ctx._h, comp.ctypes.data_as(C.POINTER(C.c_int32)), len(comp)))
        T._chk(L.tc_boruvka_phase1(self.ctx._h, self.containment, self.key.data_ptr()))
        return self.key.cpu().numpy().view(np.uint64)

    def phase2(self, comp, key):
        L, T = self.T.lib(), self.T
        self.key.copy_(self.torch.from_numpy(key.view(np.int64)))
        T._chk(L.tc_boruvka_phase2(self.ctx._h, self.containment, self.key.data_ptr(), self.edge.data_ptr()))
        return self.edge.cpu().numpy().view(np.uint64)


def boruvka(n, ops, dist_from_key, dist=None, device=None, max_rounds=64):
    """Boruvka over the union of all ranks' candidate edges.  ops.phase1(comp) -> best key per component
    (u64, U64_MAX = none); ops.phase2(comp, key) -> smallest (i<<32|j) among the edges with that key.
    Returns the forest as a structured array sorted by (dist, pre, suf) -- identical on every rank."""
    from . import EDGE_DTYPE
    parent = np.arange(n, dtype=np.int64)

    def find_all():
        p = parent
        while True:
            q = p[p]
            if np.array_equal(q, p):
                return p
            p = q

    out = []
    for _ in range(max_rounds):
        parent[:] = find_all()
        comp = parent.astype(np.int32)
        key = allreduce_min_u64(ops.phase1(comp), dist, device)
        edge = allreduce_min_u64(ops.phase2(comp, key), dist, device)
        merged = 0
        for c in np.nonzero(edge != U64_MAX)[0]:
            i, j = int(edge[c] >> np.uint64(32)), int(edge[c] & np.uint64(0xFFFFFFFF))
            a, b = i, j
            while parent[a] != a:
                a = parent[a]
            while parent[b] != b:
                b = parent[b]
            if a == b:
                continue
            parent[a] = b
            out.append((i, j, dist_from_key(int(key[c]))))
            merged += 1
        if not merged:
            break
    mst = np.array(out, dtype=EDGE_DTYPE) if out else np.zeros(0, EDGE_DTYPE)
    return mst[np.lexsort((mst["suf"], mst["pre"], mst["dist"]))]


def shard_bounds(total, rank, world, align=2):
    """[lo, hi) of the hash array owned by `rank`: equal shares rounded to `align` elements, the last rank takes
    the remainder.  Pure function (tests/test_dist_cpu.py)."""
    per = (total // world) // align * align
    lo = rank * per
    hi = total if rank == world - 1 else lo + per
    return lo, hi


def upload_sharded(ctx, hashes_pinned, off, dist, device):
    """Make the N sketches resident on EVERY rank while each rank moves only 1/world of the raw hashes over its own
    PCIe link: H2D of the rank's share, all-gather of the shares over NVLink ("one ncclAllGather of
    sketches so every GPU holds all N sketches for HP-2"), then tc_sketches_upload_dev (rank transform on the device).
    hashes_pinned: torch uint8/int64 pinned host tensor viewing all hashes (only the own share is read).
    Returns the bytes this rank copied from the host."""
    import torch
    import rabbittclust_b200 as T
    world, rank = dist.get_world_size(), dist.get_rank()
    h = hashes_pinned.view(torch.int64)
    total = int(off[-1])
    per = (total // world) // 2 * 2
    tail = total - per * world                                   # remainder, replicated by every rank (tiny)
    full = torch.empty(total + 1, dtype=torch.int64, device=device)
    lo = rank * per
    share = torch.empty(per, dtype=torch.int64, device=device)
    share.copy_(h[lo:lo + per], non_blocking=True)
    if per:
        dist.all_gather_into_tensor(full[:per * world], share)
    if tail:
        full[per * world:total].copy_(h[per * world:total], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    off = np.ascontiguousarray(off, np.uint64)
    ctx.n = len(off) - 1
    ctx.sizes = np.diff(off).astype(np.int64)
    T._chk(T.lib().tc_sketches_upload_dev(ctx._h, full.data_ptr(), T._p(off, C.c_uint64), ctx.n))
    return (per + tail) * 8


def upload_sharded_runs(ctx, hashes_pinned, off, dist, device):
    """Like upload_sharded, but no rank sorts more than its share: every rank copies 1/world of the raw hashes over its
    own PCIe link and sorts it with the positions (tc_upload_sort_share_dev); the sorted runs -- not the raw hashes --
    are all-gathered over NVLink (12 B per hash) and every rank MERGES them (tc_sketches_upload_runs_dev) instead of
    radix-sorting all N*s hashes again.  Returns the bytes this rank copied from the host."""
    import torch
    import rabbittclust_b200 as T
    L = T.lib()
    world, rank = dist.get_world_size(), dist.get_rank()
    h = hashes_pinned.view(torch.int64)
    total = int(off[-1])
    per = (total // world) // 2 * 2
    tail = total - per * world                                   # remainder: one more (tiny) run, sorted by every rank
    lo = rank * per
    share = torch.empty(max(per, 1), dtype=torch.int64, device=device)
    keys = torch.empty(max(per, 1), dtype=torch.int64, device=device)
    pos = torch.empty(max(per, 1), dtype=torch.int32, device=device)
    all_keys = torch.empty(per * world + tail + 1, dtype=torch.int64, device=device)
    all_pos = torch.empty(per * world + tail + 1, dtype=torch.int32, device=device)
    if per:
        share[:per].copy_(h[lo:lo + per], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        T._chk(L.tc_upload_sort_share_dev(ctx._h, share.data_ptr(), per, lo, keys.data_ptr(), pos.data_ptr()))
        dist.all_gather_into_tensor(all_keys[:per * world], keys[:per])
        dist.all_gather_into_tensor(all_pos[:per * world], pos[:per])
    if tail:
        t_in = torch.empty(tail, dtype=torch.int64, device=device)
        t_in.copy_(h[per * world:total], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        T._chk(L.tc_upload_sort_share_dev(ctx._h, t_in.data_ptr(), tail, per * world, all_keys[per * world:].data_ptr(),
                                          all_pos[per * world:].data_ptr()))
    torch.cuda.current_stream().synchronize()
    run_off = np.array([r * per for r in range(world + 1)] + ([total] if tail else []), np.uint64)
    if not per:
        run_off = np.array([0, total], np.uint64)
    off = np.ascontiguousarray(off, np.uint64)
    ctx.n = len(off) - 1
    ctx.sizes = np.diff(off).astype(np.int64)
    T._chk(L.tc_sketches_upload_runs_dev(ctx._h, all_keys.data_ptr(), all_pos.data_ptr(), T._p(run_off, C.c_uint64),
                                         len(run_off) - 1, T._p(off, C.c_uint64), ctx.n))
    return (per + tail) * 8


class VrRank:
    """One rank of the value-range sharded rank transform (include/tclust.h, "Multi-GPU upload, second form").  The
    methods are the per-rank steps; whoever drives them performs the exchanges in between -- upload_sharded_vr with
    torch.distributed, tests/test_upload_runs_gpu.py by hand between simulated ranks on one GPU."""
    SAMPLE = 2048

    def __init__(self, ctx, rank, world, total, device):
        import torch
        import rabbittclust_b200 as T
        self.T, self.torch, self.L = T, torch, T.lib()
        self.ctx, self.rank, self.world, self.total, self.dev = ctx, rank, world, int(total), device
        self.per = max(2, -(-self.total // world))                       # share length (the last share is shorter)
        self.lo = min(self.total, rank * self.per)
        self.len = min(self.total, self.lo + self.per) - self.lo

    def load_share(self, hashes_host):
        """H2D of the own share; returns its sample (numpy u64) for the splitters."""
        torch = self.torch
        self.share = torch.empty(max(self.len, 1), dtype=torch.int64, device=self.dev)
        if self.len:
            self.share[:self.len].copy_(hashes_host[self.lo:self.lo + self.len], non_blocking=True)
        step = max(1, self.len // self.SAMPLE)
        smp = self.share[:self.len:step][:self.SAMPLE]
        out = torch.full((self.SAMPLE,), -1, dtype=torch.int64, device=self.dev)      # all-ones = "no sample"
        out[:len(smp)] = smp
        return out

    @staticmethod
    def splitters(samples_u64, world):
        """G-1 ascending splitters from the gathered samples (numpy u64, all-ones entries = padding)."""
        v = np.sort(samples_u64[samples_u64 != U64_MAX])
        if len(v) == 0:
            return np.zeros(world - 1, np.uint64)
        return np.ascontiguousarray(v[[min(len(v) - 1, (i * len(v)) // world) for i in range(1, world)]], np.uint64)

    def partition_by_value(self, spl):
        torch, T = self.torch, self.T
        self.pk = torch.empty(max(self.len, 1), dtype=torch.int64, device=self.dev)
        self.pp = torch.empty(max(self.len, 1), dtype=torch.int32, device=self.dev)
        cnt = np.zeros(self.world, np.uint64)
        torch.cuda.current_stream().synchronize()
        T._chk(self.L.tc_vr_partition_dev(self.ctx._h, self.share.data_ptr(), None, self.len, 0, T._p(spl, C.c_uint64), 0, self.world,
                                          self.lo, self.pk.data_ptr(), self.pp.data_ptr(), T._p(cnt, C.c_uint64)))
        return cnt.astype(np.int64)

    def local_ranks(self, keys, pos):
        """keys/pos: the (hash, position) pairs of this rank's value range (device tensors).  Returns the distinct count."""
        torch, T = self.torch, self.T
        m = int(keys.numel())
        self.m = m
        self.rpos = torch.empty(max(m, 1), dtype=torch.int32, device=self.dev)
        self.rrank = torch.empty(max(m, 1), dtype=torch.int32, device=self.dev)
        nd = C.c_uint64(0)
        torch.cuda.current_stream().synchronize()
        T._chk(self.L.tc_vr_local_ranks_dev(self.ctx._h, keys.data_ptr(), pos.data_ptr(), m, self.rpos.data_ptr(), self.rrank.data_ptr(), C.byref(nd)))
        return int(nd.value)

    def partition_by_owner(self):
        torch, T = self.torch, self.T
        self.qr = torch.empty(max(self.m, 1), dtype=torch.int32, device=self.dev)
        self.qp = torch.empty(max(self.m, 1), dtype=torch.int32, device=self.dev)
        cnt = np.zeros(self.world, np.uint64)
        T._chk(self.L.tc_vr_partition32_dev(self.ctx._h, self.rrank.data_ptr(), self.rpos.data_ptr(), self.m, self.per, self.world,
                                            self.qr.data_ptr(), self.qp.data_ptr(), T._p(cnt, C.c_uint64)))
        return cnt.astype(np.int64)

    def place(self, pos, rank, seg_counts, seg_base):
        """pos/rank: the pairs that came back, grouped by source range; returns this rank's slice (length per) of the rank array."""
        torch, T = self.torch, self.T
        out = torch.zeros(self.per, dtype=torch.int32, device=self.dev)
        seg_off = np.zeros(self.world + 1, np.uint64)
        seg_off[1:] = np.cumsum(np.asarray(seg_counts, np.uint64))
        base = np.ascontiguousarray(seg_base, np.uint64)
        torch.cuda.current_stream().synchronize()
        T._chk(self.L.tc_vr_scatter_ranks_dev(self.ctx._h, pos.data_ptr(), rank.data_ptr(), T._p(seg_off, C.c_uint64), T._p(base, C.c_uint64),
                                              self.world, self.lo, self.len, out.data_ptr()))
        return out

    def finish(self, ranks_full, off):
        torch, T = self.torch, self.T
        off = np.ascontiguousarray(off, np.uint64)
        self.ctx.n = len(off) - 1
        self.ctx.sizes = np.diff(off).astype(np.int64)
        torch.cuda.current_stream().synchronize()
        T._chk(self.L.tc_sketches_upload_ranks_dev(self.ctx._h, ranks_full.data_ptr(), T._p(off, C.c_uint64), self.ctx.n))


def upload_sharded_vr(ctx, hashes_pinned, off, dist, device):
    """Multi-GPU upload with the rank transform sharded by hash value range (see VrRank): every rank copies 1/world of the
    hashes, sorts 1/world of them, and only 4-byte ranks are all-gathered.  Returns the bytes this rank copied from the host."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    total = int(off[-1])
    r = VrRank(ctx, rank, world, total, device)
    smp = r.load_share(hashes_pinned.view(torch.int64))
    allsmp = torch.empty(world * VrRank.SAMPLE, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allsmp, smp)
    spl = VrRank.splitters(allsmp.cpu().numpy().view(np.uint64), world)
    send = r.partition_by_value(spl)
    cnts = torch.empty(world * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(cnts, torch.from_numpy(send).to(device))
    cm = cnts.cpu().numpy().reshape(world, world)                     # cm[s][d]: pairs rank s sends to range d
    recv = cm[:, rank]
    m = int(recv.sum())
    keys = torch.empty(max(m, 1), dtype=torch.int64, device=device)
    pos = torch.empty(max(m, 1), dtype=torch.int32, device=device)
    dist.all_to_all_single(keys[:m], r.pk[:r.len], recv.tolist(), send.tolist())
    dist.all_to_all_single(pos[:m], r.pp[:r.len], recv.tolist(), send.tolist())
    nd = r.local_ranks(keys[:m], pos[:m])
    nds = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(nds, torch.tensor([nd], dtype=torch.int64, device=device))
    base = np.concatenate([[0], np.cumsum(nds.cpu().numpy())[:-1]])
    back = r.partition_by_owner()
    dist.all_gather_into_tensor(cnts, torch.from_numpy(back).to(device))
    bm = cnts.cpu().numpy().reshape(world, world)                     # bm[s][d]: pairs range s sends back to owner d
    rb = bm[:, rank]
    mb = int(rb.sum())
    bpos = torch.empty(max(mb, 1), dtype=torch.int32, device=device)
    brank = torch.empty(max(mb, 1), dtype=torch.int32, device=device)
    dist.all_to_all_single(bpos[:mb], r.qp[:r.m], rb.tolist(), back.tolist())
    dist.all_to_all_single(brank[:mb], r.qr[:r.m], rb.tolist(), back.tolist())
    mine = r.place(bpos, brank, rb, base)
    full = torch.empty(r.per * world, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(full, mine)
    r.finish(full, off)
    return r.len * 8


def distributed_mst(ctx, k, threshold, containment=False, dist=None):
    """Full multi-GPU MST of the sketches uploaded (replicated) on every rank's ctx."""
    import rabbittclust_b200 as T
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    ctx.allpairs_pass(T.radio(threshold, k), rank, world)
    ops = GpuOps(ctx, containment)
    L = T.lib()
    return boruvka(ctx.n, ops, lambda key: float(L.tc_dist_from_key(key, int(containment), k)), dist, ops.dev)
