"""Worker of tests/test_dist_cpu.py: one rank of a world_size-2 gloo run of the host-side multi-GPU MST logic
(rabbittclust_b200.mst_dist) with an oracle-backed stand-in for the per-rank CUDA reductions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleOps:
    """phase1/phase2 of tc_boruvka_* restated with numpy over this rank's share of the candidate edges."""

    def __init__(self, edges, sizes, containment):
        self.i, self.j, c = edges["i"].astype(np.int64), edges["j"].astype(np.int64), edges["common"].astype(np.float64)
        s0, s1 = sizes[self.i], sizes[self.j]
        denom = np.minimum(s0, s1) if containment else s0 + s1 - edges["common"]
        J = c / denom.astype(np.float64)
        self.key = ~J.view(np.uint64)
        self.id = (self.i.astype(np.uint64) << np.uint64(32)) | self.j.astype(np.uint64)

    def phase1(self, comp):
        n = len(comp)
        best = np.full(n, 0xFFFFFFFFFFFFFFFF, np.uint64)
        ci, cj = comp[self.i], comp[self.j]
        out = ci != cj
        np.minimum.at(best, ci[out], self.key[out])
        np.minimum.at(best, cj[out], self.key[out])
        return best

    def phase2(self, comp, key):
        n = len(comp)
        best = np.full(n, 0xFFFFFFFFFFFFFFFF, np.uint64)
        ci, cj = comp[self.i], comp[self.j]
        out = ci != cj
        a = out & (self.key == key[ci])
        b = out & (self.key == key[cj])
        np.minimum.at(best, ci[a], self.id[a])
        np.minimum.at(best, cj[b], self.id[b])
        return best


def main(rank, world, port, containment, out_dir):
    import torch.distributed as dist
    import rabbittclust_b200 as T
    from rabbittclust_b200 import mst_dist, synth
    from oracle import pyoracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k, thr = 21, 0.05
    if containment:
        sk = synth.family_sketches(6, 15, seed=21, ragged=(100, 1500))
    else:
        sk = synth.family_sketches(7, 14, s=600, seed=20)
    n = len(sk)
    sizes = np.array([len(s) for s in sk], np.int64)
    cm = O.common_matrix(sk)
    radio = T.radio(thr, k)
    # this rank's share of the reference candidate graph (tile ownership rule of the dense pass)
    rows = []
    for i in range(n):
        for j in range(i):
            r, c = i // 32, j // 32
            if mst_dist.tile_owner(r * (r + 1) // 2 + c, world) != rank:
                continue
            mn, mx = min(sizes[i], sizes[j]), max(sizes[i], sizes[j])
            if cm[i, j] >= 1 and mn > 0 and not mx > radio * mn:
                rows.append((i, j, cm[i, j]))
    edges = np.array(rows, dtype=T.PAIR_DTYPE)
    L = T.lib()
    mst = mst_dist.boruvka(n, OracleOps(edges, sizes, containment),
                           lambda key: float(L.tc_dist_from_key(key, int(containment), k)), dist)
    pairs = mst_dist.owned_pairs(n, rank, world)
    np.save(os.path.join(out_dir, "mst_%d.npy" % rank), mst)
    np.save(os.path.join(out_dir, "pairs_%d.npy" % rank), np.array([pairs, len(edges)]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
