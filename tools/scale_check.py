#!/usr/bin/env python
"""Full-size runs of the BASELINE.json parity-case shapes that are not bench lines (configs[3] and configs[4]),
checked through size-independent properties (the oracle cannot run these sizes in reasonable time).

  kssd   configs[4] shape: N KSSD-like sketches (about 488 u32-range hashes in a 2^24 space, k = 18) -> tc_mst.
         Properties: N-1-(components-1) forest edges sorted by weight; the partition at d equals the planted
         families (cross-family pairs share only chance hashes: J << Jmin); every in-family MST edge weight equals
         the host formula on the exact intersection recomputed with numpy.
  greedy configs[3] shape: N ragged sketches (100..10000, containment, k = 21) sorted by length desc ->
         tc_greedy_cluster.  Properties: every member is within d of its representative (exact host distance on the
         numpy intersection) and every representative is farther than d from every earlier representative of its
         own family sample; cluster count reported.
usage: python tools/scale_check.py kssd|greedy [N] [--gpus G]
--gpus G: additionally the same job on G GPUs of the box -- kssd: one context + host thread per device in a multi-GPU
group (tc_mg_*: work units of the pass dealt round-robin, Boruvka reductions over NVLink peer memory inside the kernel);
greedy: tc_greedy_cluster_multi (representative table dealt round-robin) -- and the result must equal the 1-GPU one.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rabbittclust_b200 as T  # noqa: E402


def kssd_like(n_fam, per_fam, size=488, space=1 << 24, seed=5):
    rng = np.random.default_rng(seed)
    sk = []
    for f in range(n_fam):
        pool = rng.choice(space, size, replace=False).astype(np.uint64)
        for m in range(per_fam):
            keep = int(size * rng.uniform(0.6, 1.0))
            part = rng.choice(pool, keep, replace=False)
            priv = rng.integers(0, space, size - keep + 4).astype(np.uint64)
            s = np.unique(np.concatenate([part, priv]))
            sk.append(s)
    return sk


def ragged_families(n_fam, per_fam, seed=6, lo=100, hi=10000):
    """member = the s smallest values of (random p-subset of the family pool + private values)"""
    rng = np.random.default_rng(seed)
    sk, lens = [], []
    for f in range(n_fam):
        pool = np.sort(rng.integers(0, 2**62, hi, dtype=np.uint64)) * np.uint64(2)
        for m in range(per_fam):
            s = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
            p = rng.uniform(0.3, 1.0)
            top = pool[:s]
            part = top[rng.random(s) < p]
            priv = rng.integers(0, int(top[-1]) // 2 + 1, s - len(part), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
            v = np.unique(np.concatenate([part, priv]))
            sk.append(v)
            lens.append(1000 * len(v) + int(rng.integers(0, 1000)))
    return sk, np.asarray(lens)


def group_mst(sk_flat, off, k, d, gpus):
    """(upload_s, mst_s, mst) of a G-GPU group: threads of this process, one context per device."""
    import threading
    ctxs = [T.Context(r) for r in range(gpus)]
    n = len(off) - 1
    handles = [c.mg_export(n) for c in ctxs]
    out, err, t_up, t_mst = [None] * gpus, [None] * gpus, [0.0] * gpus, [0.0] * gpus
    bar = threading.Barrier(gpus)

    def work(r):
        try:
            ctxs[r].mg_join(r, gpus, handles)
            bar.wait()
            t0 = time.time(); ctxs[r].upload(hashes=sk_flat, off=off); t_up[r] = time.time() - t0
            bar.wait()
            ctxs[r].mst(k, d)                                  # warm-up (unit lists, buffers)
            bar.wait()
            t0 = time.time(); out[r] = ctxs[r].mst(k, d).copy(); t_mst[r] = time.time() - t0
        except Exception as ex:      # noqa: BLE001
            bar.abort(); err[r] = ex
    th = [threading.Thread(target=work, args=(r,)) for r in range(gpus)]
    [t.start() for t in th]
    [t.join() for t in th]
    stats = [c.mst_stats() for c in ctxs]
    for c in ctxs:
        c.mg_leave(); c.close()
    for e in err:
        if e is not None:
            raise e
    for r in range(1, gpus):
        assert np.array_equal(out[0], out[r]), "ranks disagree"
    return max(t_up), max(t_mst), out[0], stats


def main():
    gpus = 1
    if "--gpus" in sys.argv:
        i = sys.argv.index("--gpus"); gpus = int(sys.argv[i + 1]); del sys.argv[i:i + 2]
    what = sys.argv[1]
    res = {"what": what}
    ctx = T.Context(0)
    if what == "kssd":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
        per = 100
        t0 = time.time(); sk = kssd_like(n // per, per); res["gen_s"] = time.time() - t0
        n = len(sk)
        k, d = 18, 0.05
        flat, off = T.pack_sketches(sk)
        t0 = time.time(); ctx.upload(hashes=flat, off=off); res["upload_s"] = time.time() - t0
        t0 = time.time(); mst = ctx.mst(k, d); res["mst_first_s"] = time.time() - t0       # may include the edge-buffer growth
        t0 = time.time(); mst = ctx.mst(k, d).copy(); res["mst_s"] = time.time() - t0
        res["stats"] = ctx.mst_stats()
        res.update(n=n, pairs=n * (n - 1) // 2, mst_edges=len(mst), pairs_per_s_incl_boruvka=n * (n - 1) / 2 / res["mst_s"])
        if gpus > 1:
            ctx.close()
            up, ms, m2, st = group_mst(flat, off, k, d, gpus)
            res["multi_gpu"] = {"gpus": gpus, "upload_s": up, "mst_s": ms, "speedup_vs_1gpu": res["mst_s"] / ms,
                                "equals_1gpu": bool(np.array_equal(m2, mst)),
                                "pass_ms_per_rank": [x["pass_ms"] for x in st], "boruvka_ms_per_rank": [x["boruvka_ms"] for x in st],
                                "edges_per_rank": [x["edges"] for x in st]}
            assert res["multi_gpu"]["equals_1gpu"]
            ctx = T.Context(0)
        assert (np.diff(mst["dist"]) >= 0).all()
        # union-find over edges <= d
        par = np.arange(n)
        def find(x):
            while par[x] != x:
                par[x] = par[par[x]]; x = par[x]
            return x
        for e in mst[mst["dist"] <= d]:
            a, b = find(int(e["pre"])), find(int(e["suf"]))
            if a != b:
                par[a] = b
        roots = np.array([find(i) for i in range(n)])
        fam = np.arange(n) // per
        # each cluster lies inside one family and each family is one cluster (members share >= 60 % of the pool)
        assert len(set(zip(roots.tolist(), fam.tolist()))) == len(set(roots.tolist())) == n // per, "partition != families"
        rng = np.random.default_rng(0)
        for t in rng.integers(0, len(mst), 300):
            i, j = int(mst["pre"][t]), int(mst["suf"][t])
            c = len(np.intersect1d(sk[i], sk[j], assume_unique=True))
            assert mst["dist"][t] == T.mash_dist(c, len(sk[i]), len(sk[j]), k)
        res["clusters"] = n // per
    else:
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
        per = 100
        t0 = time.time(); sk, lens = ragged_families(n // per, per); res["gen_s"] = time.time() - t0
        order = np.lexsort((np.arange(len(sk)), -lens))          # length desc, id asc (the --presketched order)
        sk = [sk[i] for i in order]
        fam = (order // per)
        n = len(sk)
        k, d = 21, 0.05
        t0 = time.time(); ctx.upload(sk); res["upload_s"] = time.time() - t0
        t0 = time.time(); nc, rep = ctx.greedy_cluster(k, d, containment=True); res["greedy_s"] = time.time() - t0
        res.update(n=n, clusters=nc, mean_size=float(np.mean([len(s) for s in sk])))
        if gpus > 1:
            import threading
            flat, off = T.pack_sketches(sk)
            ctxs = [ctx] + [T.Context(r) for r in range(1, gpus)]
            th = [threading.Thread(target=lambda c=c: c.upload(hashes=flat, off=off)) for c in ctxs[1:]]
            t0 = time.time(); [t.start() for t in th]; [t.join() for t in th]; up = time.time() - t0
            t0 = time.time(); nc2, rep2 = T.greedy_cluster_multi(ctxs, k, d, containment=True); gs = time.time() - t0
            for c in ctxs[1:]:
                c.close()
            res["multi_gpu"] = {"gpus": gpus, "upload_s": up, "greedy_s": gs, "speedup_vs_1gpu": res["greedy_s"] / gs,
                                "equals_1gpu": bool(nc2 == nc and np.array_equal(rep2, rep))}
            assert res["multi_gpu"]["equals_1gpu"]
        assert nc == len(set(rep.tolist())) and (rep <= np.arange(n)).all() and (rep[rep] == rep).all()
        L = T.lib()
        rng = np.random.default_rng(1)
        for j in rng.integers(0, n, 2000):
            j = int(j); r = int(rep[j])
            if r == j:
                continue
            c = len(np.intersect1d(sk[j], sk[r], assume_unique=True))
            assert L.tc_greedy_dist(c, len(sk[j]), len(sk[r]), 1, k) <= d, (j, r)
        # representatives of the same family must not be within d of an earlier representative
        reps = np.nonzero(rep == np.arange(n))[0]
        checked = 0
        for f in rng.integers(0, n // per, 40):
            rf = reps[fam[reps] == f]
            for a in range(len(rf)):
                for b in range(a):
                    i, j = int(rf[a]), int(rf[b])
                    c = len(np.intersect1d(sk[i], sk[j], assume_unique=True))
                    cmin = L.tc_greedy_common_min(d, k, len(sk[i]), len(sk[j]), 1)
                    assert c < cmin or L.tc_greedy_dist(c, len(sk[i]), len(sk[j]), 1, k) > d, (i, j)
                    checked += 1
        res["rep_pairs_checked"] = checked
    ctx.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
