"""tc_upload_sort_share_dev + tc_sketches_upload_runs_dev (the multi-GPU upload: every rank sorts its share, all ranks
merge the gathered runs) on one GPU: the runs of 1, 2, 3 and 5 "ranks" give the context the plain upload gives."""
import ctypes as C

import numpy as np
import pytest

from rabbittclust_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shares", [1, 2, 3, 5])
def test_upload_from_sorted_runs(tc, shares):
    import torch
    sk = synth.family_sketches(6, 14, s=300, seed=51) + synth.family_sketches(3, 8, seed=52, ragged=(40, 900))
    sk.append(np.zeros(0, np.uint64))
    flat, off = tc.pack_sketches(sk)
    total, n = int(off[-1]), len(sk)
    L = tc.lib()
    dev = torch.device("cuda", 0)
    with tc.Context(0) as ref:
        ref.upload(sk)
        want = ref.pair_common(0, n, 0, n, impl=2)
        want_mst = ref.mst(21, 0.05).copy()
    with tc.Context(0) as c:
        h = torch.from_numpy(flat[:total].view(np.int64)).to(dev)
        keys = torch.empty(total + 1, dtype=torch.int64, device=dev)
        pos = torch.empty(total + 1, dtype=torch.int32, device=dev)
        bounds = [total * r // shares for r in range(shares + 1)]
        for r in range(shares):
            lo, hi = bounds[r], bounds[r + 1]
            tc._chk(L.tc_upload_sort_share_dev(c._h, h[lo:].data_ptr(), hi - lo, lo, keys[lo:].data_ptr(), pos[lo:].data_ptr()))
        torch.cuda.synchronize()
        k = keys[:total].cpu().numpy().view(np.uint64)
        p = pos[:total].cpu().numpy()
        for r in range(shares):
            lo, hi = bounds[r], bounds[r + 1]
            assert np.all(np.diff(k[lo:hi].astype(np.float64)) >= 0) and np.array_equal(np.sort(p[lo:hi]), np.arange(lo, hi))
            assert np.array_equal(flat[p[lo:hi]], k[lo:hi])
        run_off = np.array(bounds, np.uint64)
        c.n, c.sizes = n, np.diff(off).astype(np.int64)
        tc._chk(L.tc_sketches_upload_runs_dev(c._h, keys.data_ptr(), pos.data_ptr(), tc._p(run_off, C.c_uint64), shares,
                                              tc._p(off, C.c_uint64), n))
        assert np.array_equal(c.pair_common(0, n, 0, n, impl=2), want)
        assert np.array_equal(c.pair_common(0, n, 0, n, impl=0), want)
        assert np.array_equal(c.mst(21, 0.05), want_mst)
        with pytest.raises(tc.TclustError, match="original hashes"):
            c.pair_common(0, 2, 0, 2, impl=1)
        # an unsorted sketch is still rejected (the check runs on the ranks)
        bad = flat.copy()
        bad[[2, 3]] = bad[[3, 2]]
        hb = torch.from_numpy(bad[:total].view(np.int64)).to(dev)
        tc._chk(L.tc_upload_sort_share_dev(c._h, hb.data_ptr(), total, 0, keys.data_ptr(), pos.data_ptr()))
        one = np.array([0, total], np.uint64)
        rc = L.tc_sketches_upload_runs_dev(c._h, keys.data_ptr(), pos.data_ptr(), tc._p(one, C.c_uint64), 1, tc._p(off, C.c_uint64), n)
        assert rc != 0 and b"sketch 0 is not strictly ascending" in L.tc_last_error()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_value_range_sharded_transform(tc, world):
    """The value-range sharded rank transform (mst_dist.VrRank) with `world` simulated ranks on one GPU, the exchanges done
    by hand: the resident set equals the one of the plain upload (pair counts, MST, edges)."""
    import torch
    from rabbittclust_b200.mst_dist import VrRank
    sk = synth.family_sketches(7, 12, s=250, seed=61) + synth.family_sketches(3, 9, seed=62, ragged=(30, 700))
    sk.insert(5, np.zeros(0, np.uint64))
    flat, off = tc.pack_sketches(sk)
    total, n = int(off[-1]), len(sk)
    dev = torch.device("cuda", 0)
    with tc.Context(0) as ref:
        ref.upload(sk)
        want = ref.pair_common(0, n, 0, n, impl=2)
        want_mst = ref.mst(21, 0.05).copy()
    hosts = torch.from_numpy(flat[:total].view(np.int64).copy())
    ctxs = [tc.Context(0) for _ in range(world)]
    try:
        rs = [VrRank(c, r, world, total, dev) for r, c in enumerate(ctxs)]
        smp = torch.cat([r.load_share(hosts) for r in rs]).cpu().numpy().view(np.uint64)
        spl = VrRank.splitters(smp, world)
        send = np.stack([r.partition_by_value(spl) for r in rs])             # send[s][d]
        nds = []
        for d, r in enumerate(rs):                                              # all-to-all by hand
            ks, ps = [], []
            for s_, src in enumerate(rs):
                o = int(send[s_, :d].sum())
                ks.append(src.pk[o:o + int(send[s_, d])]); ps.append(src.pp[o:o + int(send[s_, d])])
            nds.append(r.local_ranks(torch.cat(ks).contiguous(), torch.cat(ps).contiguous()))
        base = np.concatenate([[0], np.cumsum(nds)[:-1]])
        back = np.stack([r.partition_by_owner() for r in rs])                  # back[s][d]
        slices = []
        for d, r in enumerate(rs):
            ps, rk = [], []
            for s_, src in enumerate(rs):
                o = int(back[s_, :d].sum())
                ps.append(src.qp[o:o + int(back[s_, d])]); rk.append(src.qr[o:o + int(back[s_, d])])
            slices.append(r.place(torch.cat(ps).contiguous(), torch.cat(rk).contiguous(), back[:, d], base))
        full = torch.cat(slices).contiguous()
        for r in rs:
            r.finish(full, off)
        for c in ctxs:
            assert np.array_equal(c.pair_common(0, n, 0, n, impl=2), want)
            assert np.array_equal(c.mst(21, 0.05), want_mst)
        # the ranks are the dense ranks of the plain transform: compare through the union of all hashes
        ranks = (full[:total].cpu().numpy().view(np.uint32) & np.uint32(0x7FFFFFFF)).astype(np.int64)
        uniq, inv, cnt = np.unique(flat[:total], return_inverse=True, return_counts=True)
        assert np.array_equal(ranks, inv)
        assert np.array_equal((full[:total].cpu().numpy().view(np.uint32) >> np.uint32(31)).astype(bool), cnt[inv] > 1)
    finally:
        for c in ctxs:
            c.close()
