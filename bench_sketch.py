"""bench_sketch.py -- HP-1 leg of bench.py: MinHash sketch construction (K1) on synthetic genomes of
the BASELINE.json configs[1] shape (5 Mbp, single record, k=21, s=1000), generated ON THE DEVICE
(50,000 x 5 Mbp does not fit host RAM/disk, so a step sketches one HBM-resident batch).

value = genomes/s with the batch resident in HBM (CUDA events on the library's stream);
e2e   = genomes/s through tc_sketch_minhash with HOST (pinned) sequence buffers: H2D of the bases,
        kernels, D2H of the sketches inside the timed region.
"""
import ctypes as C
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def device_genomes(torch, n, length, seed, device):
    """n genomes of `length` iid uniform ACGT bases as one uint8 device tensor (+64 B tail pad)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    total = n * length
    buf = torch.zeros(total + 64, dtype=torch.uint8, device=device)
    step = 1 << 28
    for o in range(0, total, step):
        m = min(step, total - o)
        c = torch.randint(0, 4, (m,), dtype=torch.uint8, device=device, generator=g)
        # 0,1,2,3 -> 'A','C','G','T' (65,67,71,84)
        buf[o:o + m] = 65 + 2 * c + 2 * (c >= 2).to(torch.uint8) + 11 * (c == 3).to(torch.uint8)
        del c
    return buf


def measure(args, local_rank=0, rank=0, world=1, dist=None):
    """Returns a dict with the sketch-leg numbers of this rank's batch (genomes are independent units:
    every rank sketches its own batch, no collective -> weak scaling)."""
    import torch
    import rabbittclust_b200 as T
    L = T.lib()
    dev = torch.device("cuda", local_rank)
    ng, glen, k, s = args.genomes, args.genome_len, 21, 1000
    ctx = T.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    seq = device_genomes(torch, ng, glen, 1234 + rank, dev)
    rec_off = (np.arange(ng + 1, dtype=np.uint64) * np.uint64(glen))
    r2g = np.arange(ng, dtype=np.uint32)
    ss = np.full(ng, s, np.uint32)
    out = torch.zeros((ng, s), dtype=torch.int64, device=dev)
    out_n = torch.zeros(ng, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step():
        T._chk(L.tc_sketch_minhash_dev(ctx._h, seq.data_ptr(), T._p(rec_off, C.c_uint64), T._p(r2g, C.c_uint32),
                                       ng, ng, k, T._p(ss, C.c_uint32), 42, out.data_ptr(), s, out_n.data_ptr()))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(args.steps):
        step()
    with torch.cuda.stream(stream):
        e1.record()
    e1.synchronize()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launches - l0
    # size-independent property at full size: every sketch has exactly s strictly ascending hashes
    o = out.cpu().numpy().view(np.uint64)
    assert (out_n.cpu().numpy() == s).all(), "sketch count"
    assert (o[:, 1:] > o[:, :-1]).all(), "sketches not strictly ascending"

    # ---- e2e: host pinned bases -> tc_sketch_minhash -> host sketches
    ne = max(1, min(ng, args.e2e_genomes))
    h_seq = torch.empty(ne * glen + 64, dtype=torch.uint8).pin_memory()
    h_seq[:ne * glen].copy_(seq[:ne * glen])
    h_seq[ne * glen:] = 0
    h_out = torch.zeros((ne, s), dtype=torch.int64).pin_memory()
    h_n = torch.zeros(ne, dtype=torch.int32).pin_memory()
    ro, rg, sz = rec_off[:ne + 1].copy(), r2g[:ne].copy(), ss[:ne].copy()

    def e2e_step():
        T._chk(L.tc_sketch_minhash(ctx._h, h_seq.data_ptr(), T._p(ro, C.c_uint64), T._p(rg, C.c_uint32),
                                   ne, ne, k, T._p(sz, C.c_uint32), 42, h_out.data_ptr(), s, h_n.data_ptr()))

    e2e_step()
    torch.cuda.synchronize()
    n_e2e = max(1, min(args.steps, 3))
    t0 = time.time()
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.time() - t0) * 1e3 / n_e2e
    assert np.array_equal(h_out.numpy().view(np.uint64), o[:ne]), "host-buffer path differs from device path"
    pcie_bytes = int(L.tc_last_h2d_bytes(ctx._h))      # what the last call moved over PCIe (packed 2-bit batches count packed)
    ctx.close()
    del seq
    return {"genomes": ng, "genome_len": glen, "k": k, "s": s, "ms_per_step": dev_ms,
            "genomes_per_s": ng / (dev_ms / 1e3), "gbp_per_s": ng * glen / (dev_ms / 1e3) / 1e9,
            "launches": int(launches), "e2e_genomes": ne, "e2e_ms_per_step": e2e_ms,
            "e2e_genomes_per_s": ne / (e2e_ms / 1e3), "h2d_bytes": ne * glen, "d2h_bytes": ne * s * 8 + ne * 4,
            "pcie_bytes": pcie_bytes}


def measure_kssd(args, local_rank=0, rank=0):
    """KSSD (--fast) sketch leg, BASELINE configs[4] shape: 2 Mbp genomes, k=18 (half_k 9), drlevel 3."""
    import torch
    import rabbittclust_b200 as T
    L = T.lib()
    dev = torch.device("cuda", local_rank)
    ng, glen, k, drl, stride = args.genomes, 2000000, 18, 3, 2048
    ctx = T.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    seq = device_genomes(torch, ng, glen, 4321 + rank, dev)
    rec_off = (np.arange(ng + 1, dtype=np.uint64) * np.uint64(glen))
    r2g = np.arange(ng, dtype=np.uint32)
    sd = T.kssd_shuffled_dim(6)
    out = torch.zeros((ng, stride), dtype=torch.int64, device=dev)
    out_n = torch.zeros(ng, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step():
        T._chk(L.tc_sketch_kssd_dev(ctx._h, seq.data_ptr(), T._p(rec_off, C.c_uint64), T._p(r2g, C.c_uint32), ng, ng, k, drl,
                                    T._p(sd, C.c_int32), out.data_ptr(), stride, out_n.data_ptr()))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(args.steps):
        step()
    with torch.cuda.stream(stream):
        e1.record()
    e1.synchronize()
    dev_ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launches - l0
    n = out_n.cpu().numpy()
    o = out.cpu().numpy().view(np.uint64)
    assert (n > 300).all() and (n < 700).all(), "KSSD sketch sizes out of the expected range"   # ~ L / 4096 = 488
    for g in range(0, ng, max(1, ng // 64)):
        assert (o[g, 1:n[g]] > o[g, :n[g] - 1]).all(), "KSSD sketch not strictly ascending"
    ne = max(1, min(ng, args.e2e_genomes // 2))
    h_seq = torch.empty(ne * glen + 64, dtype=torch.uint8).pin_memory()
    h_seq[:ne * glen].copy_(seq[:ne * glen]); h_seq[ne * glen:] = 0
    h_out = torch.zeros((ne, stride), dtype=torch.int64).pin_memory()
    h_n = torch.zeros(ne, dtype=torch.int32).pin_memory()
    ro, rg = rec_off[:ne + 1].copy(), r2g[:ne].copy()

    def e2e_step():
        T._chk(L.tc_sketch_kssd(ctx._h, h_seq.data_ptr(), T._p(ro, C.c_uint64), T._p(rg, C.c_uint32), ne, ne, k, drl,
                                T._p(sd, C.c_int32), h_out.data_ptr(), stride, h_n.data_ptr()))

    e2e_step()
    torch.cuda.synchronize()
    n_e2e = max(1, min(args.steps, 3))
    t0 = time.time()
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.time() - t0) * 1e3 / n_e2e
    assert np.array_equal(h_n.numpy(), n[:ne]) and np.array_equal(h_out.numpy().view(np.uint64)[:, :256], o[:ne, :256])
    ctx.close()
    return {"metric": "genomes sketched/sec (KSSD --fast)", "value": ng / (dev_ms / 1e3), "unit": "genomes/s",
            "ms_per_step": dev_ms, "gbp_per_s": ng * glen / (dev_ms / 1e3) / 1e9, "gpu_launches": int(launches),
            "config": {"workload": "KSSD sketch: %d synthetic %d bp genomes, k=18 drlevel=3 (BASELINE.json configs[4] shape)"
                                   % (ng, glen), "mean_sketch_size": float(n.mean())},
            "e2e": {"value": ne / (e2e_ms / 1e3), "unit": "genomes/s", "h2d_bytes_per_step": ne * glen,
                    "d2h_bytes_per_step": ne * stride * 8 + ne * 4, "genomes_per_step": ne, "ms_per_step": e2e_ms}}


def cpu_update(args, threads):
    """The reference's MinHash::update (oracle/_ref/ref_harness, AVX2 build) on a bounded sample."""
    import subprocess
    from oracle import pyoracle
    exe = pyoracle.ref_bin("ref_harness")
    if exe is None:
        return None
    ng = max(threads, 8)
    r = json.loads(subprocess.check_output([exe, "update", "21", "1000", str(threads), str(args.genome_len), str(ng)]))
    return r


def sketch_block(args, m, world, clocks=None, with_cpu=True):
    """The sketch-leg numbers as a JSON-able dict (same keys as a bench line)."""
    from bench import measured_peak
    peak, how = measured_peak()
    ng, glen, s = m["genomes"], m["genome_len"], m["s"]
    value = world * ng / (m["ms_per_step"] / 1e3)
    alg_bytes = ng * (glen + 8 * s)
    achieved = alg_bytes / (m["ms_per_step"] / 1e3) / 1e9
    blk = {"metric": "genomes sketched/sec", "value": value, "unit": "genomes/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": "sketch-only: %d synthetic %d bp genomes per GPU and step, k=21 s=1000 "
                                  "(BASELINE.json configs[1] shape, device-generated)" % (ng, glen),
                      "l2": "input %.1f GB per step > 126 MB L2" % (ng * glen / 1e9),
                      "gbp_per_s": world * m["gbp_per_s"]},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "peak_source": how, "kernel": "k1_sketch_segment_kernel",
                        "note": "algorithmic bytes = L + 8s per genome (SURVEY 8d); K1 is integer-issue bound "
                                "(Murmur3 = ten 64-bit multiplies per window), not HBM bound: see DESIGN.md K1"},
           "e2e": {"value": world * m["e2e_genomes_per_s"], "unit": "genomes/s",
                   "h2d_bytes_per_step": m["h2d_bytes"] * world, "d2h_bytes_per_step": m["d2h_bytes"] * world,
                   "genomes_per_step": m["e2e_genomes"] * world, "ms_per_step": m["e2e_ms_per_step"],
                   "pcie_bytes_per_step": m.get("pcie_bytes", m["h2d_bytes"]) * world,
                   "transport": ("host ASCII buffers; the library packs each batch to 2 bits per base on the host threads "
                                 "(+ exception list), the device rebuilds the bytes" if m.get("pcie_bytes", m["h2d_bytes"]) < m["h2d_bytes"]
                                 else "host ASCII buffers copied as they are (1 byte per base)")},
           "gpu_launches": m["launches"]}
    if clocks is not None:
        blk["clocks"] = clocks
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            prof = json.load(open(tp))
            blk["roofline"]["traffic"] = prof.get("dram_bytes_per_launch_at_bench_size")
            blk["roofline"]["binding_units"] = {k: prof[k] for k in prof if k.endswith("_pct")}
        except Exception:
            pass
    if with_cpu:
        threads = os.cpu_count() or 1
        r = cpu_update(args, threads)
        if r:
            blk["cpu_baseline"] = {"value": r["genomes_per_s"], "unit": "genomes/s", "cores": r["threads"],
                                   "kind": "reference", "sample": "%d genomes of %d bp, MinHash::update AVX2 build"
                                   % (r["genomes"], r["len"])}
    return blk


def run_sketch(args):
    """--workload sketch: the sketch leg as the main JSON line."""
    import torch
    from bench import ClockSampler, dist_env
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    if args.workload == "kssd":
        from bench import emit
        emit(measure_kssd(args, local_rank, rank))
        return
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m = measure(args, local_rank, rank, world, dist)
    clocks = sampler.stop() if rank == 0 else None
    m = reduce_ranks(torch, dist, m)
    if rank == 0:
        from bench import emit
        emit(sketch_block(args, m, world, clocks, with_cpu=(world == 1 and not args.no_cpu)))
    if dist is not None:
        dist.destroy_process_group()


def reduce_ranks(torch, dist, m):
    """max of the times over the ranks, sum of the launches"""
    if dist is None:
        return m
    t = torch.tensor([m["ms_per_step"], m["e2e_ms_per_step"], float(m["launches"])], dtype=torch.float64, device="cuda")
    tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    m = dict(m)
    m["ms_per_step"], m["e2e_ms_per_step"], m["launches"] = float(tm[0]), float(tm[1]), int(ts[2])
    m["e2e_genomes_per_s"] = m["e2e_genomes"] / (m["e2e_ms_per_step"] / 1e3)
    m["genomes_per_s"] = m["genomes"] / (m["ms_per_step"] / 1e3)
    m["gbp_per_s"] = m["genomes"] * m["genome_len"] / (m["ms_per_step"] / 1e3) / 1e9
    return m
