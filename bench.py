#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot paths (contract: task brief, section 4).

Headline (default, --workload pairs): ONE dense all-pairs pass (K3J block-join kernel over the per-block union
lists, units where it does not pay on the K3H hash-join kernel; candidate-edge epilogue + Boruvka round 1;
TCLUST_PASS_IMPL=join|merge selects K3H / the first-generation merge kernel) over N=100,000 pre-sketched s=1000
MinHash sketches (BASELINE.json configs[2], the configuration the north-star's pairs/s target is quoted on; it fits
one GPU).
    value  = pairwise sketch distances / s, sketches resident in HBM (CUDA events)
    e2e    = same through the C ABI with HOST buffers: H2D upload + rank transform + union lists + pass + D2H edges
--workload sketch: MinHash sketch construction, synthetic 5 Mbp genomes k=21 s=1000 (configs[1]).
--impl reference: the reference's own CPU path -- the unmodified clust-mst binary (oracle/_ref, AVX2 build) on the
full workload, its inverted-index generateMST phase; MinHash::distance brute force on a bounded sample beside it --
all host threads.

Multi-GPU: one process per GPU (torchrun); sketches replicated, the tile space of the fixed
N=100k problem is dealt round-robin to the ranks ("strong" scaling); per step one
all-reduce(min) pair of the Boruvka round-1 keys over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: libraries that print to fd 1 (e.g. NCCL's "NCCL version ..." banner under
# NCCL_DEBUG=VERSION) are redirected to stderr for the whole run; emit() writes the line to the real stdout
if "TCLUST_BENCH_STDOUT_FD" not in os.environ or os.environ.get("TCLUST_BENCH_STDOUT_PID") != str(os.getpid()):
    os.environ["TCLUST_BENCH_STDOUT_FD"] = str(os.dup(1))      # (bench_sketch imports this module a second time)
    os.environ["TCLUST_BENCH_STDOUT_PID"] = str(os.getpid())
    os.dup2(2, 1)
_REAL_STDOUT = int(os.environ["TCLUST_BENCH_STDOUT_FD"])


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


PAIR_BYTES = 16000.0   # algorithmic operand bytes per pair: 8*(|A|+|B|) at s=1000


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason sampling during the timed region.  The timed region of the pairs leg is tens of
    milliseconds, so the samples come from NVML directly (a thread polling every ~2 ms; the ctypes/torch calls of the
    timed loop release the GIL); `nvidia-smi -lms` is the fallback when pynvml cannot be used."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.period_ms = int(os.environ.get("TCLUST_BENCH_SAMPLER_MS", "2"))
        self.nvml, self.handle, self.thread, self.stop_flag = None, None, None, threading.Event()
        self.samples, self.reason_bits, self.max_mhz = [], 0, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:
            import torch
            u = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
        except Exception:
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def _poll(self):
        nv, h = self.nvml, self.handle
        while True:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            except Exception:
                pass
            try:
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                try:
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                except Exception:
                    pass
            if self.stop_flag.wait(self.period_ms / 1e3):
                return

    def start(self):
        if self.period_ms <= 0:
            return
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(max(self.period_ms, 20))],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv, b = self.nvml, self.reason_bits
            names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                     ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                     ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                     ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown)]
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(self.samples), "reasons": sorted(k for k, m in names if b & m), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons, "source": "nvidia-smi"}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------- reference arm
def ref_pairs(n_sample, threads, hashes, off, k=21, s=1000):
    """Time the reference's MinHash::distance all-pairs (oracle/_ref/ref_harness) on n_sample sketches."""
    import tempfile
    from rabbittclust_b200 import sketch_io
    from oracle import pyoracle
    exe = pyoracle.ref_bin("ref_harness")
    if exe is None:
        return None
    with tempfile.TemporaryDirectory() as d:
        sk = [hashes[int(off[i]):int(off[i + 1])] for i in range(n_sample)]
        sketch_io.write_minhash_folder(d, sk, ["g%06d" % i for i in range(n_sample)], k=k, sketch_size=s)
        out = subprocess.check_output([exe, "pairs", os.path.join(d, "hash.sketch"), str(threads), str(n_sample)])
    return json.loads(out)


def ref_index_mst(n, threads, hashes, off, runs, workdir=None, k=21, s=1000, keep=None):
    """The reference's own default path for this workload: the unmodified clust-mst binary (oracle/_ref) on a
    --presketched folder of the n sketches -- compute_minhash_mst with its inverted index (src/MST.cpp:1290-1737), all
    host threads.  The first run builds and caches the index (minhash.sketch.index) and is reported separately; the
    following `runs` runs are timed through the binary's own "generateMST" phase line.  None without the binary."""
    import re
    import shutil
    import tempfile
    from rabbittclust_b200 import sketch_io
    from oracle import pyoracle
    exe = pyoracle.ref_bin("clust-mst")
    if exe is None or not np.all(np.diff(off[:n + 1]) == s):
        return None
    d = tempfile.mkdtemp(prefix="tclust_ref_", dir=workdir)
    try:
        folder = os.path.join(d, "sketches")
        sketch_io.write_minhash_folder_uniform(folder, hashes, n, s, k)
        res = []
        for it in range(runs + 1):
            t0 = time.time()
            r = subprocess.run([exe, "-t", str(threads), "--presketched", folder, "-d", "0.05", "-o", "out.txt"], cwd=d,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            wall = time.time() - t0
            if r.returncode:
                return None
            ph = {m.group(1).strip(): float(m.group(2))
                  for m in re.finditer(r"=+time of ([^=]+?) is: ([0-9.eE+-]+)", r.stderr.decode())}
            if "generateMST" not in ph:
                return None
            res.append({"generate_mst_s": ph["generateMST"], "load_s": ph.get("load genome Infos and sketch Infos"),
                        "wall_s": wall})
        return {"pairs": n * (n - 1) // 2, "threads": threads, "first_run": res[0], "runs": res[1:]}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def pyoracle_ref_harness():
    from oracle import pyoracle
    return pyoracle.ref_bin("ref_harness")


def port_pairs(n_sample, hashes, off):
    """Scalar oracle port (full-intersection merge), single thread."""
    from oracle import pyoracle
    sk = [hashes[int(off[i]):int(off[i + 1])] for i in range(n_sample)]
    t0 = time.time()
    pyoracle.common_matrix(sk)
    dt = time.time() - t0
    return {"pairs": n_sample * n_sample, "seconds": dt, "pairs_per_s": n_sample * n_sample / dt, "threads": 1}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from rabbittclust_b200 import synth
    threads = os.cpu_count() or 1
    if args.workload == "pairs":
        # (1) the reference's default path on the FULL workload: clust-mst --presketched, generateMST phase
        idx = None
        if not args.ref_bruteforce_only:
            hf, offf = synth.family_sketches_fast(args.n // 100, 100, 1000, seed=7)
            idx = ref_index_mst(len(offf) - 1, threads, hf, offf, args.steps)
            del hf
        # (2) MinHash::distance (the AVX2 kernel north_star names) over a bounded sample
        n_sample = args.ref_sample
        h, off = synth.family_sketches_fast(max(n_sample // 100, 1), 100, 1000, seed=7)
        n_sample = len(off) - 1
        vals = []
        kind = "reference"
        for it in range((args.warmup + args.steps) if idx is None else 2):
            r = ref_pairs(n_sample, threads, h, off)
            if r is None:
                r = port_pairs(min(n_sample, 600), h, off)
                kind = "port"
            if it >= (args.warmup if idx is None else 1):
                vals.append(r)
        pairs = sum(v["pairs"] for v in vals)
        secs = sum(v["seconds"] for v in vals)
        brute = pairs / secs
        brute_block = {"value": brute, "unit": "pairs/s", "cores": vals[0].get("threads", 1), "kind": kind,
                       "what": "MinHash::distance (AVX2 build), all pairs of a bounded sample",
                       "sample": "%d sketches of the same synthetic model = %d pairs per step" % (n_sample, vals[0]["pairs"])}
        if idx is not None:
            # the stronger baseline is the headline of this arm: every step is one run of the reference binary on the
            # full workload with its index cached (the first, index-building run is the warm-up and reported beside)
            secs_i = sum(r["generate_mst_s"] for r in idx["runs"])
            value = idx["pairs"] * len(idx["runs"]) / secs_i
            ms_step = 1e3 * secs_i / len(idx["runs"])
            cpu = {"value": value, "unit": "pairs/s", "cores": idx["threads"], "kind": "reference",
                   "what": "unmodified clust-mst --presketched: compute_minhash_mst with its inverted index, generateMST phase",
                   "sample": "the full workload (%d sketches, %d pairs) per step; index cached by the first run" %
                             (args.n, idx["pairs"]),
                   "first_run_generate_mst_s": idx["first_run"]["generate_mst_s"],
                   "generate_mst_s_steps": [round(r["generate_mst_s"], 3) for r in idx["runs"]],
                   "wall_s_steps": [round(r["wall_s"], 3) for r in idx["runs"]],
                   "minhash_distance": brute_block}
            note = ("CPU arm: the reference's own default path (inverted-index compute_minhash_mst inside the unmodified "
                    "binary) on the full workload; warm-up = the one index-building run; MinHash::distance brute force "
                    "beside it in cpu_baseline.minhash_distance")
        else:
            value, ms_step, cpu = brute, 1e3 * secs / len(vals), brute_block
            note = "CPU arm: MinHash::distance (AVX2 build) over a bounded sample"
        line = {"impl": "reference", "metric": "pairwise sketch distances/sec", "value": value, "unit": "pairs/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": "clust-mst all-pairs on 100,000 pre-sketched s=1000 k=21 sketches "
                                       "(BASELINE.json configs[2])", "note": note},
                "cpu_baseline": cpu,
                "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        # the sketch half of the metric on the same arm: MinHash::update of the reference on all host threads
        exe = pyoracle_ref_harness()
        if exe and not args.no_sketch:
            ng = max(threads, 8)
            r = json.loads(subprocess.check_output([exe, "update", "21", "1000", str(threads), str(args.genome_len), str(ng)]))
            line["sketch"] = {"metric": "genomes sketched/sec", "value": r["genomes_per_s"], "unit": "genomes/s",
                              "cpu_baseline": {"value": r["genomes_per_s"], "unit": "genomes/s", "cores": r["threads"],
                                               "kind": "reference", "sample": "%d genomes of %d bp, MinHash::update AVX2 build"
                                               % (r["genomes"], r["len"])},
                              "e2e": {"value": r["genomes_per_s"], "unit": "genomes/s", "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": 0}}
    else:
        from oracle import pyoracle
        exe = pyoracle.ref_bin("ref_harness")
        ng, L = max(threads, 8), args.genome_len
        vals, kind = [], "reference"
        for it in range(args.warmup + args.steps):
            if exe:
                r = json.loads(subprocess.check_output([exe, "update", "21", "1000", str(threads), str(L), str(ng)]))
            else:
                kind = "port"
                rng = np.random.default_rng(1)
                g = synth.random_genome(L, rng)
                t0 = time.time(); pyoracle.minhash_sketch([g], 21, 1000); dt = time.time() - t0
                r = {"genomes": 1, "seconds": dt, "threads": 1}
            if it >= args.warmup:
                vals.append(r)
        secs = sum(v["seconds"] for v in vals)
        value = sum(v["genomes"] for v in vals) / secs
        line = {"impl": "reference", "metric": "genomes sketched/sec", "value": value, "unit": "genomes/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * secs / len(vals), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": "sketch-only synthetic %d bp genomes k=21 s=1000 (BASELINE.json configs[1])" % L},
                "cpu_baseline": {"value": value, "unit": "genomes/s", "cores": vals[0].get("threads", 1), "kind": kind,
                                 "sample": "%d genomes of %d bp per step" % (vals[0]["genomes"], L)},
                "e2e": {"value": value, "unit": "genomes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------ our arm
def run_pairs(args):
    import torch
    import rabbittclust_b200 as T
    from rabbittclust_b200 import synth
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.n
    per_fam = 100
    h, off = synth.family_sketches_fast(max(n // per_fam, 1), per_fam, 1000, seed=7)
    n = len(off) - 1
    if args.order == "shuffled":
        # same sketches in a random order: no two neighbouring sketches are related, so a row block of 32 sketches
        # holds 32,000 distinct hashes (3 table partitions) instead of ~9,000 (1 partition) -- see DESIGN.md K3H
        perm = np.random.default_rng(11).permutation(n)
        h = np.ascontiguousarray(h.reshape(n, 1000)[perm].reshape(-1))
    k, thr = 21, 0.05
    radio = T.radio(thr, k)
    ctx = T.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    h_pin = torch.from_numpy(h).pin_memory()
    hp = h_pin.numpy()
    best_key = torch.empty(n, dtype=torch.int64, device="cuda")
    best_edge = torch.empty(n, dtype=torch.int64, device="cuda")
    comp = np.arange(n, dtype=np.int32)
    L = T.lib()
    import ctypes as C

    # The Boruvka phases of a step are queued back to back on the library's stream (tc_boruvka_set_async): phase 1,
    # all-reduce(min) of the keys, phase 2, all-reduce(min) of the edges -- NCCL orders itself after the current stream,
    # which is the library's here -- and the host waits once, at the end of the round.
    T._chk(L.tc_boruvka_set_async(ctx._h, 1))

    def boruvka_round1():
        with torch.cuda.stream(stream):
            T._chk(L.tc_boruvka_upload_comp(ctx._h, comp.ctypes.data_as(C.POINTER(C.c_int32)), n))
            T._chk(L.tc_boruvka_phase1(ctx._h, 0, best_key.data_ptr()))
            if dist is not None:
                # keys are u64 bit patterns < 2^63 is NOT guaranteed: reduce as unsigned via the sign flip
                best_key.bitwise_xor_(-2**63); dist.all_reduce(best_key, op=dist.ReduceOp.MIN); best_key.bitwise_xor_(-2**63)
            T._chk(L.tc_boruvka_phase2(ctx._h, 0, best_key.data_ptr(), best_edge.data_ptr()))
            if dist is not None:
                best_edge.bitwise_xor_(-2**63); dist.all_reduce(best_edge, op=dist.ReduceOp.MIN); best_edge.bitwise_xor_(-2**63)
        # a Boruvka round ends with the reduced edges on every rank (the host unions the components next):
        # the phases and collectives are complete before the next pass starts, they are not left to overlap it
        stream.synchronize()

    def device_step():
        npairs, nedges = ctx.allpairs_pass(radio, rank, world)
        boruvka_round1()
        return npairs, nedges

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.upload(hashes=hp, off=off)
    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync_all()        # after the sampler's start-up (NVML init, a few ms on rank 0): the other ranks must not wait for it inside their timed region
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k3_ms, pairs_mine, nedges, k3_steps, kern_steps = 0.0, 0, 0, [], []
    with torch.cuda.stream(stream):
        ev0.record()
    t_wall0 = time.time()
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record()
        npairs, nedges = ctx.allpairs_pass(radio, rank, world)
        with torch.cuda.stream(stream):
            b.record()
        b.synchronize()
        k3_steps.append(a.elapsed_time(b))            # whole tc_allpairs_pass call (unit upload, kernel, count read-back)
        kern_steps.append(ctx.last_pass_kernel_ms)    # events recorded by the library right around the kernel
        k3_ms += kern_steps[-1]
        pairs_mine += npairs
        boruvka_round1()                                   # round complete on this rank when it returns
    with torch.cuda.stream(stream):
        ev1.record()
    sync_all()
    ev1.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    wall_ms = (time.time() - t_wall0) * 1e3
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers -> upload (H2D + rank transform) -> pass -> edges D2H
    e2e_steps = max(1, min(args.steps, 3))
    ctx.edges(pinned=True)               # the page-locked result buffer exists before the timed steps (it is reused)
    sync_all()
    t0 = time.time()
    d2h = 0
    h2d = int(h.nbytes + off.nbytes)
    for _ in range(e2e_steps):
        if dist is None:
            ctx.upload(hashes=hp, off=off)
        else:
            # every rank copies only its 1/world share from the host; the shares are all-gathered over NVLink
            from rabbittclust_b200 import mst_dist
            h2d = mst_dist.upload_sharded(ctx, h_pin, off, dist, torch.device("cuda", local_rank)) + int(off.nbytes)
        ctx.allpairs_pass(radio, rank, world)
        e = ctx.edges(pinned=True)
        d2h = e.nbytes
    sync_all()
    e2e_ms = (time.time() - t0) * 1e3 / e2e_steps

    # ---- semantic B (MinHash::distance, the computation the CPU arm times): one pass of the step-budget merge kernel
    semb_ms, semb_pairs = None, 0
    if not args.no_semb:
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record()
        semb_pairs, _ = ctx.allpairs_pass_b(1000, rank, world)
        with torch.cuda.stream(stream):
            b.record()
        b.synchronize()
        semb_ms = a.elapsed_time(b)
        sync_all()

    t = torch.tensor([dev_ms, e2e_ms, k3_ms, float(pairs_mine), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms = float(tmax[0]), float(tmax[1])
        total_pairs, launches = float(tsum[3]), int(tsum[4])
        gl = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(gl, torch.tensor([k3_ms / args.steps], dtype=torch.float64, device="cuda"))
        per_rank_ms = [float(x[0]) for x in gl]
    else:
        total_pairs = float(pairs_mine)
        per_rank_ms = [k3_ms / args.steps]
    if rank == 0:
        peak, how = measured_peak()
        pairs_per_step_all = n * (n - 1) / 2
        value = total_pairs / (dev_ms / 1e3)
        k3_avg_s = k3_ms / 1e3 / args.steps
        achieved = (pairs_mine / args.steps) * PAIR_BYTES / k3_avg_s / 1e9
        traffic, binding = None, None
        impl_env = os.environ.get("TCLUST_PASS_IMPL")
        kernel = ("k3_tile_kernel" if impl_env == "merge" else "k3h_join_kernel" if impl_env == "join"
                  else "k3j_block_kernel (units where the block join does not pay run on k3h_join_kernel)")
        tp = os.path.join(ROOT, "profiles", "k3_traffic.json")
        if os.path.exists(tp) and impl_env is None and args.order == "family" and n == 100000 and world == 1:
            try:
                prof = json.load(open(tp))
                traffic = prof.get("dram_bytes_per_launch")
                binding = {k: prof[k] for k in ("issue_active_pct", "alu_pipe_pct", "lsu_pipe_pct",
                                                "shared_wavefronts_pct", "l2_hit_pct") if k in prof}
                binding["source"] = "ncu capture of this command, profiles/k3_traffic.json"
            except Exception:
                traffic = None
        line = {
            "metric": "pairwise sketch distances/sec", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 ranks of u64 hashes",
            "data": "synthetic",
            "config": {"workload": "clust-mst dense all-pairs pass on %d pre-sketched s=1000 k=21 sketches, "
                                   "1000-per-family model (BASELINE.json configs[2])" % n,
                       "n_sketches": n, "sketch_size": 1000, "pairs_per_step": pairs_per_step_all,
                       "order": args.order + (" (generation order: the 100 members of a family are adjacent)"
                                              if args.order == "family" else " (random permutation of the same sketches)"),
                       "candidate_edges": int(nedges), "parallelism": "tile-sharded x%d, sketches replicated" % world,
                       "l2": "input 0.4 GB of u32 ranks (+0.8 GB u64) per GPU > 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                         "kernel": kernel, "kernel_ms": 1e3 * k3_avg_s, "kernel_ms_per_rank": per_rank_ms,
                         "kernel_ms_steps": [round(x, 3) for x in kern_steps],
                         "pass_call_ms_steps": [round(x, 3) for x in k3_steps],
                         "binding_units": binding,
                         "note": "algorithmic bytes = SURVEY 8(d)'s operand bytes 8*(|A|+|B|) = 16000 B/pair x pairs per "
                                 "launch.  Every distinct rank of a 32-sketch column block is probed once per 32 row "
                                 "sketches (shared-memory hash join over per-block union lists), so the operand bytes "
                                 "are served from shared memory/L2, not HBM: frac > 1 is expected, `traffic` is the "
                                 "measured DRAM bytes, and the binding units are the issue slots / ALU pipe / "
                                 "shared-memory pipe (DESIGN.md K3J)"},
            "e2e": {"value": (pairs_per_step_all) / (e2e_ms / 1e3), "unit": "pairs/s",
                    "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
        }
        if semb_ms is not None:
            line["semantic_b"] = {"what": "the same pass with Mash union-limited counts (MinHash::jaccard stop rule, "
                                          "sketchSize 1000) on the step-budget merge kernel k3_tile_kernel<EDGES,SEMB>; "
                                          "this is the computation the cpu_baseline runs; rank 0's share",
                                  "value": semb_pairs / (semb_ms / 1e3), "unit": "pairs/s", "ms": semb_ms,
                                  "pairs": int(semb_pairs)}
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            r = ref_pairs(min(args.ref_sample, n), threads, h, off)
            kind = "reference"
            if r is None:
                r, kind = port_pairs(min(600, n), h, off), "port"
            brute = {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r.get("threads", 1),
                     "kind": kind, "sample": "all pairs of the first %d sketches (%d pairs), "
                     "MinHash::distance AVX2 build" % (min(args.ref_sample, n), r["pairs"])}
            idx = None
            if args.order == "family" and not args.ref_bruteforce_only:
                idx = ref_index_mst(n, threads, h, off, 1)
            if idx is not None:
                line["cpu_baseline"] = {
                    "value": idx["pairs"] / idx["runs"][0]["generate_mst_s"], "unit": "pairs/s", "cores": idx["threads"],
                    "kind": "reference",
                    "sample": "the full workload once: unmodified clust-mst --presketched on the same %d sketches, "
                              "generateMST phase (inverted-index compute_minhash_mst), index cached by a first run" % n,
                    "generate_mst_s": idx["runs"][0]["generate_mst_s"], "wall_s": idx["runs"][0]["wall_s"],
                    "first_run_generate_mst_s": idx["first_run"]["generate_mst_s"],
                    "minhash_distance": brute}
            else:
                line["cpu_baseline"] = brute
    ctx.close()
    # ---- second half of BASELINE.json's metric: genomes sketched/sec (configs[1] shape), same JSON line
    if not args.no_sketch:
        import bench_sketch
        sargs = argparse.Namespace(**vars(args))
        m = bench_sketch.reduce_ranks(torch, dist, bench_sketch.measure(sargs, local_rank, rank, world, dist))
        if rank == 0:
            line["sketch"] = bench_sketch.sketch_block(sargs, m, world, None, with_cpu=(world == 1 and not args.no_cpu))
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pairs", choices=["pairs", "sketch", "kssd"])
    ap.add_argument("--n", type=int, default=100000, help="number of sketches (pairs workload)")
    ap.add_argument("--genomes", type=int, default=2048, help="genomes per step (sketch workload)")
    ap.add_argument("--genome-len", type=int, default=5000000)
    ap.add_argument("--e2e-genomes", type=int, default=512, help="genomes per e2e step (sketch workload)")
    ap.add_argument("--ref-sample", type=int, default=4000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-bruteforce-only", action="store_true",
                    help="CPU legs: only MinHash::distance over the sample, not the reference binary on the full workload")
    ap.add_argument("--no-sketch", action="store_true", help="pairs workload: skip the sketch leg of the line")
    ap.add_argument("--no-semb", action="store_true", help="pairs workload: skip the semantic-B pass")
    ap.add_argument("--order", default="family", choices=["family", "shuffled"],
                    help="pairs workload: sketch order (family = as generated, SURVEY 8d config 3; shuffled = random)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "pairs":
        return run_pairs(args)
    from bench_sketch import run_sketch
    return run_sketch(args)


if __name__ == "__main__":
    main()
