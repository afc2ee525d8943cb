"""Worker of tests/test_dist_gpu.py (torchrun, one rank per GPU, NCCL): multi-GPU MST of a replicated sketch
set; every rank writes its MST; rank 0 also writes the oracle's MST for comparison."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_dir, containment):
    import torch
    import torch.distributed as dist
    import rabbittclust_b200 as T
    from rabbittclust_b200 import mst_dist, synth
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sk = (synth.family_sketches(8, 40, seed=41, ragged=(100, 2500)) if containment
          else synth.family_sketches(10, 40, s=1000, seed=40))
    ctx = T.Context(local)
    if containment:
        ctx.upload(sk)
    else:
        # sharded upload: every rank reads only its share of the pinned host hashes, NCCL all-gathers the rest
        h, off = T.pack_sketches(sk)
        h_pin = torch.from_numpy(h).pin_memory()
        mst_dist.upload_sharded(ctx, h_pin, off, dist, torch.device("cuda", local))
        want = np.array([[len(np.intersect1d(sk[i], sk[j])) for j in range(3)] for i in range(37, 41)], np.uint32)
        assert np.array_equal(ctx.pair_common(37, 41, 0, 3, impl=2), want), "sharded upload differs"
    mst = mst_dist.distributed_mst(ctx, 21, 0.05, bool(containment), dist)
    np.save(os.path.join(out_dir, "mst_%d.npy" % rank), mst)
    if rank == 0:
        from oracle import pyoracle as O
        np.save(os.path.join(out_dir, "ref.npy"), O.mst(sk, 21, 0.05, bool(containment)))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
