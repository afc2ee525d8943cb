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
