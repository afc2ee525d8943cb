"""Multi-GPU (NCCL) MST: tile-sharded dense pass + all-reduce(MIN) Boruvka on >= 2 GPUs of one box equals the
oracle's MST on every rank.  Skipped on a single-GPU box (the CPU gloo test covers the host logic)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("containment", [0, 1])
def test_nccl_distributed_mst(tmp_path, containment):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(HERE, "_dist_gpu_worker.py"), str(tmp_path), str(containment)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    ref = np.load(tmp_path / "ref.npy")
    for rank in range(world):
        m = np.load(tmp_path / ("mst_%d.npy" % rank))
        assert len(m) == len(ref)
        assert np.array_equal(np.sort(m["dist"]), np.sort(ref["dist"]))
