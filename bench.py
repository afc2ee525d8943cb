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


def workload_name(n):
    """config.workload, identical in both arms (the driver compares the strings)."""
    return ("clust-mst --presketched: all-pairs sketch intersection + MST on %d pre-sketched s=1000 k=21 MinHash sketches, "
            "%d families x 100 (BASELINE.json configs[2])" % (n, n // 100))


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
        result = None
        if keep:
            # the reference's results on this folder: the cluster file the binary has just written, and its MST --
            # clust-mst --presketched never writes edge.mst, so the unmodified compute_minhash_mst is called once more
            # through ref_harness (same objects, same index cache) to get the weights
            try:
                cl = sketch_io.parse_cluster_file(os.path.join(d, "out.txt"))
                result = {"labels": labels_from_clusters(n, cl), "dist": None}
                hx = pyoracle.ref_bin("ref_harness")
                if hx is not None:
                    ob = os.path.join(d, "ref_mst.bin")
                    subprocess.run([hx, "msti", folder, str(threads), "0.05", ob], cwd=d, check=True,
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
                    result["dist"] = np.asarray(sketch_io.read_edge_mst(ob)["dist"], np.float64)
            except Exception as ex:      # the timing stands; parity is then reported as missing
                sys.stderr.write("bench: could not read the reference's result: %r\n" % (ex,))
        return {"pairs": n * (n - 1) // 2, "threads": threads, "first_run": res[0], "runs": res[1:], "result": result}
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
                "config": {"workload": workload_name(args.n // 100 * 100), "note": note},
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
def partition_labels(n, pre, suf, dist, thr):
    """Connected components of the forest edges with dist <= thr (generateForest + BFS of the reference,
    src/MST.cpp:77-142): label = smallest vertex id of the component."""
    parent = np.arange(n, dtype=np.int64)
    keep = dist <= thr
    for a, b in zip(pre[keep].tolist(), suf[keep].tolist()):
        while parent[a] != a:
            parent[a] = parent[parent[a]]; a = parent[a]
        while parent[b] != b:
            parent[b] = parent[parent[b]]; b = parent[b]
        if a != b:
            parent[max(a, b)] = min(a, b)
    while True:
        q = parent[parent]
        if np.array_equal(q, parent):
            return parent
        parent = q


def labels_from_clusters(n, clusters):
    lab = np.full(n, -1, np.int64)
    for c in clusters:
        ids = [m[0] if isinstance(m, tuple) else m for m in c]
        lab[ids] = min(ids)
    return lab


def run_pairs(args):
    import torch
    import rabbittclust_b200 as T
    from rabbittclust_b200 import synth
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    per_fam = 100
    h, off = synth.family_sketches_fast(max(n // per_fam, 1), per_fam, 1000, seed=7)
    n = len(off) - 1
    if args.order == "shuffled":
        perm = np.random.default_rng(11).permutation(n)
        h = np.ascontiguousarray(h.reshape(n, 1000)[perm].reshape(-1))
    k, thr = 21, 0.05
    ctx = T.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    h_pin = torch.from_numpy(h).pin_memory()
    hp = h_pin.numpy()

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if dist is not None:
        # multi-GPU group: every rank exports its exchange buffer, the handles travel over torch.distributed (plumbing),
        # and from then on ctx.mst() is a collective whose reductions run inside the Boruvka kernel over NVLink
        mine = torch.frombuffer(bytearray(ctx.mg_export(n)), dtype=torch.uint8).to(dev)
        allh = torch.empty(world * T.MG_HANDLE_BYTES, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine)
        ctx.mg_join(rank, world, bytes(allh.cpu().numpy().tobytes()))
        sync_all()

    def upload():
        if dist is None:
            ctx.upload(hashes=hp, off=off)
            return int(h.nbytes + off.nbytes)
        # every rank copies and sorts only its 1/world share; the sorted runs are all-gathered over NVLink and merged
        from rabbittclust_b200 import mst_dist
        if os.environ.get("TCLUST_BENCH_UPLOAD") == "allgather_raw":      # round 1's scheme: all-gather the raw hashes, every rank sorts all
            return mst_dist.upload_sharded(ctx, h_pin, off, dist, dev) + int(off.nbytes)
        if os.environ.get("TCLUST_BENCH_UPLOAD") == "vr":                 # rank transform sharded by hash value range (measured: not faster up to 8 GPUs)
            return mst_dist.upload_sharded_vr(ctx, h_pin, off, dist, dev) + int(off.nbytes)
        return mst_dist.upload_sharded_runs(ctx, h_pin, off, dist, dev) + int(off.nbytes)

    # ---- value: the whole operation on resident sketches -- dense pass (this rank's units) + every Boruvka round +
    #      the sorted MST on the host (tc_mst), CUDA events on the library's stream
    upload()
    mst_out = np.empty(n, T.EDGE_DTYPE)      # the caller's result array of tc_mst, reused from step to step as a C caller would
    for _ in range(args.warmup):
        ctx.mst(k, thr, out=mst_out)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync_all()        # after the sampler's start-up (NVML init on rank 0): no rank waits for it inside its timed region
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_steps, bor_steps, sort_steps, call_steps = [], [], [], []
    with torch.cuda.stream(stream):
        ev0.record()
    t_wall0 = time.time()
    mst = None
    for _ in range(args.steps):
        t_c = time.time()
        mst = ctx.mst(k, thr, out=mst_out)
        call_steps.append((time.time() - t_c) * 1e3)
        st = ctx.mst_stats()
        pass_steps.append(st["pass_ms"]); bor_steps.append(st["boruvka_ms"]); sort_steps.append(st["sort_ms"])
    with torch.cuda.stream(stream):
        ev1.record()
    sync_all()
    ev1.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    wall_ms = (time.time() - t_wall0) * 1e3
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    stats = ctx.mst_stats()

    # ---- e2e: host buffers -> upload (H2D + rank transform + union lists + tile flags) -> tc_mst -> MST edges on the host
    e2e_steps = max(1, min(args.steps, 3))
    sync_all()
    t0 = time.time()
    h2d = 0
    for _ in range(e2e_steps):
        h2d = upload()
        mst = ctx.mst(k, thr, out=mst_out)
    sync_all()
    e2e_ms = (time.time() - t0) * 1e3 / e2e_steps
    d2h = int(mst.nbytes)
    mst = mst.copy()          # the legs below compare against it; mst_out belongs to the timed loops

    # ---- threshold-only clustering: the clusters at d straight from the pass, no edges, no MST
    thr_only = None
    if not args.no_robust:
        sync_all()
        ctx.cluster_threshold(k, thr, False, rank, world)
        t_c = time.time()
        nc_t, lab_t = ctx.cluster_threshold(k, thr, False, rank, world)
        thr_ms = (time.time() - t_c) * 1e3
        if dist is not None:                     # the parts' labels are merged on the host (tc_labels_merge)
            allab = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(world)]
            dist.all_gather(allab, torch.from_numpy(lab_t.copy()).to(dev))
            merged = allab[0].cpu().numpy()
            for x in allab[1:]:
                nc_t, merged = T.labels_merge(merged, x.cpu().numpy())
            lab_t = merged
        if rank == 0:
            want_lab = partition_labels(n, mst["pre"], mst["suf"], mst["dist"], thr)
            thr_only = {"what": "tc_cluster_threshold: connected components of the candidate pairs within the threshold (device "
                                "union-find epilogue of the pass, O(n) memory); this rank's part of the pass",
                        "ms": round(thr_ms, 3), "clusters": int(len(np.unique(lab_t))),
                        "equals_mst_partition": bool(np.array_equal(lab_t, want_lab))}

    # ---- a set large enough to scale: one species of --dense-n sketches (every pair is a candidate, every tile dense).
    #      configs[2] itself is ~1 ms of work for one GPU, of which only the dense pass shards; this leg shows the same
    #      collective tc_mst where the sharded part dominates.  Same code path, every rank, device-timed like `value`.
    large = None
    if not args.no_large:
        hd, od = synth.family_sketches_fast(1, args.dense_n, 1000, seed=5)
        nd = len(od) - 1
        if dist is not None:
            ctx.mg_leave()
            sync_all()                     # every rank has unmapped its peers before any exchange buffer is re-created
            mine = torch.frombuffer(bytearray(ctx.mg_export(max(n, nd))), dtype=torch.uint8).to(dev)
            allh = torch.empty(world * T.MG_HANDLE_BYTES, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine)
            ctx.mg_join(rank, world, bytes(allh.cpu().numpy().tobytes()))
        ctx.upload(hashes=hd, off=od)
        ctx.mst(k, thr)
        sync_all()
        la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            la.record()
        for _ in range(3):
            md = ctx.mst(k, thr)
        with torch.cuda.stream(stream):
            lb.record()
        sync_all()
        lb.synchronize()
        sl = ctx.mst_stats()
        lt = torch.tensor([la.elapsed_time(lb) / 3, sl["pass_ms"], sl["boruvka_ms"]], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        large = {"what": "collective tc_mst on one species of %d sketches (s=1000): %d pairs, all of them candidates" % (nd, nd * (nd - 1) // 2),
                 "n": nd, "pairs": nd * (nd - 1) // 2, "ms_per_step": float(lt[0]), "value": nd * (nd - 1) / 2 / (float(lt[0]) / 1e3),
                 "unit": "pairs/s", "pass_kernel_ms_max_rank": float(lt[1]), "boruvka_kernel_ms_max_rank": float(lt[2]),
                 "mst_edges": int(len(md)), "n_gpus": world}
        del hd
        h2d_restore = upload()          # the headline set is resident again for what follows

    # ---- N > 1: the collective result must be the single-GPU result, edge for edge
    same_as_single = None
    if dist is not None and rank == 0:
        with T.Context(local_rank) as c1:
            c1.upload(hashes=hp, off=off)
            m1 = c1.mst(k, thr)
        same_as_single = bool(len(m1) == len(mst) and np.array_equal(m1["pre"], mst["pre"]) and
                              np.array_equal(m1["suf"], mst["suf"]) and np.array_equal(m1["dist"], mst["dist"]))

    # ---- semantic B (MinHash::distance, the computation the brute-force CPU leg times): one pass of the merge kernel
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

    # ---- robustness (1 GPU): the same sketches in random order, a set in which every pair is related, and the default
    #      set with the tile flags switched off (what the pass cost before them)
    robust = None
    if world == 1 and not args.no_robust:
        robust = {}

        def leg(hh, oo, env=None):
            for kk, vv in (env or {}).items():
                os.environ[kk] = vv
            try:
                with T.Context(local_rank) as c2:
                    t_u = time.time()
                    c2.upload(hashes=hh, off=oo)
                    up_ms = (time.time() - t_u) * 1e3
                    c2.mst(k, thr)
                    t_m = time.time()
                    m2 = c2.mst(k, thr)
                    mst_ms = (time.time() - t_m) * 1e3
                    s2 = c2.mst_stats()
            finally:
                for kk in (env or {}):
                    os.environ.pop(kk, None)
            nn = len(oo) - 1
            return {"n": nn, "pairs": nn * (nn - 1) // 2, "upload_ms": round(up_ms, 3), "mst_ms": round(mst_ms, 3),
                    "pass_kernel_ms": round(s2["pass_ms"], 3), "boruvka_kernel_ms": round(s2["boruvka_ms"], 3),
                    "rounds": s2["rounds"], "candidate_edges": s2["edges"], "tiles_visited": s2["tiles_flagged"],
                    "tiles_total": s2["tiles_total"], "on_reordered_copy": s2["on_reordered_copy"],
                    "mst_edges": int(len(m2))}, m2
        other = "shuffled" if args.order == "family" else "family"
        if other == "shuffled":
            perm = np.random.default_rng(11).permutation(n)
            h2 = np.ascontiguousarray(h.reshape(n, 1000)[perm].reshape(-1))
            robust["shuffled_order"], m_sh = leg(h2, off)
            # same sketches, other ids: the weights must be the same multiset
            robust["shuffled_order"]["weights_equal_family_order"] = bool(np.array_equal(np.sort(m_sh["dist"]), np.sort(mst["dist"])))
            del h2
        hd, od = synth.family_sketches_fast(1, args.dense_n, 1000, seed=5)
        robust["one_species"], _ = leg(hd, od)
        robust["one_species"]["what"] = ("one family of %d sketches: every pair shares hashes, every tile is visited and is dense "
                                         "(worst case for the tile flags)" % args.dense_n)
        robust["every_tile_visited"], m_off = leg(hp, off, {"TCLUST_TILE_FLAGS": "0"})
        robust["every_tile_visited"]["what"] = "the default set with the tile flags switched off (TCLUST_TILE_FLAGS=0): all 4.9e6 tiles streamed"
        robust["every_tile_visited"]["same_mst"] = bool(np.array_equal(m_off, mst))

    t = torch.tensor([dev_ms, e2e_ms, float(launches), float(np.mean(pass_steps)), float(np.mean(bor_steps))],
                     dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms = float(tmax[0]), float(tmax[1])
        launches = int(tsum[2])
        gl = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(gl, t[3:5].clone())
        per_rank_pass = [float(x[0]) for x in gl]
        per_rank_bor = [float(x[1]) for x in gl]
    else:
        per_rank_pass, per_rank_bor = [float(np.mean(pass_steps))], [float(np.mean(bor_steps))]
    line = None
    if rank == 0:
        peak, how = measured_peak()
        pairs_per_step = n * (n - 1) / 2
        value = pairs_per_step * args.steps / (dev_ms / 1e3)
        pass_s = max(per_rank_pass) / 1e3
        achieved = (pairs_per_step / world) * PAIR_BYTES / pass_s / 1e9
        prof, traffic, binding = None, None, None
        tp = os.path.join(ROOT, "profiles", "k3_traffic.json")
        impl_env = os.environ.get("TCLUST_PASS_IMPL")
        if os.path.exists(tp) and impl_env is None and args.order == "family" and n == 100000 and world == 1:
            try:
                prof = json.load(open(tp))
                traffic = prof.get("dram_bytes_per_launch")
                binding = {kk: prof[kk] for kk in ("issue_active_pct", "alu_pipe_pct", "lsu_pipe_pct",
                                                   "shared_wavefronts_pct", "l2_hit_pct", "kernel_ms_under_ncu") if kk in prof}
                binding["source"] = "committed ncu capture of this command (profiles/k3_traffic.json), not measured in this run"
            except Exception:
                prof = None
        kernel = ("k3_tile_kernel" if impl_env == "merge" else "k3h_join_kernel" if impl_env == "join"
                  else "k3j_block_kernel" if impl_env == "blockjoin"
                  else "k3s_tile_kernel (flagged tiles only, union lists staged with TMA bulk copies; row blocks without a union "
                       "list run on k3h_join_kernel)")
        line = {
            "metric": "pairwise sketch distances/sec", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 ranks of u64 hashes",
            "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "step": "tc_mst on resident sketches: dense all-pairs pass (candidate edges) + all Boruvka rounds on "
                               "the device + sorted MST edges read back to the host",
                       "n_sketches": n, "sketch_size": 1000, "pairs_per_step": pairs_per_step,
                       "order": args.order + (" (generation order: the 100 members of a family are adjacent)"
                                              if args.order == "family" else " (random permutation of the same sketches)"),
                       "candidate_edges_rank0": stats["edges"], "boruvka_rounds": stats["rounds"],
                       "tiles_visited": stats["tiles_flagged"], "tiles_total": stats["tiles_total"],
                       "mst_edges": int(len(mst)),
                       "parallelism": "work units of the pass dealt round-robin x%d, sketches replicated, Boruvka "
                                      "min-reduction fused into the kernel over NVLink peer memory" % world,
                       "l2": "input 0.4 GB of u32 ranks (+0.8 GB u64) per GPU > 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                         "kernel": kernel, "kernel_ms": 1e3 * pass_s, "kernel_ms_per_rank": per_rank_pass,
                         "kernel_ms_steps": [round(x, 3) for x in pass_steps],
                         "boruvka_kernel_ms_per_rank": per_rank_bor,
                         "boruvka_kernel_ms_steps": [round(x, 3) for x in bor_steps],
                         "forest_sort_ms_steps": [round(x, 3) for x in sort_steps],
                         "tc_mst_call_ms_steps": [round(x, 3) for x in call_steps],
                         "issue_frac": (prof or {}).get("issue_active_pct", None) and prof["issue_active_pct"] / 100.0,
                         "dram_frac": (traffic / pass_s / 1e9 / peak) if traffic else None,
                         "binding_units": binding,
                         "note": "`achieved`/`frac` follow SURVEY 8(d): operand bytes 8*(|A|+|B|) = 16000 B/pair x the pairs "
                                 "this launch covers / kernel time.  It is NOT a utilisation: the kernel only visits the tiles "
                                 "whose sketch blocks share a rank (config.tiles_visited of config.tiles_total) and serves "
                                 "those from shared memory/L2, so it moves a tiny fraction of the operand bytes.  What it "
                                 "really uses of the chip: dram_frac = measured DRAM bytes (ncu) / kernel time / HBM peak, "
                                 "issue_frac = issue slots busy (ncu), both from the committed capture"},
            "e2e": {"value": pairs_per_step / (e2e_ms / 1e3), "unit": "pairs/s",
                    "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                    "ms_per_step": e2e_ms,
                    "what": "host hashes -> tc_sketches_upload -> tc_mst -> sorted MST edges on the host (all Boruvka rounds)",
                    "pcie": {"h2d_gb_per_rank": round(h2d / 1e9, 3),
                             "copy_ms_at_55_gbs": round(h2d / 55e9 * 1e3, 2),
                             "note": "the host->device copy of the u64 hashes is the floor of this step: at the ~55 GB/s a "
                                     "PCIe Gen5 x16 link sustains it alone takes copy_ms_at_55_gbs of ms_per_step"}},
            "gpu_launches": int(launches), "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
        }
        if same_as_single is not None:
            line["mst_equals_single_gpu"] = same_as_single
        if robust is not None:
            line["robustness"] = robust
        if large is not None:
            line["large_set"] = large
        if thr_only is not None:
            line["threshold_only"] = thr_only
        if semb_ms is not None:
            line["semantic_b"] = {"what": "the same pass with Mash union-limited counts (MinHash::jaccard stop rule, "
                                          "sketchSize 1000) on the step-budget merge kernel k3_tile_kernel<EDGES,SEMB>; "
                                          "this is the computation cpu_baseline.minhash_distance runs; rank 0's share",
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
            if not args.ref_bruteforce_only:
                idx = ref_index_mst(n, threads, h, off, 1, keep=True)
            if idx is not None:
                line["cpu_baseline"] = {
                    "value": idx["pairs"] / idx["runs"][0]["generate_mst_s"], "unit": "pairs/s", "cores": idx["threads"],
                    "kind": "reference",
                    "sample": "the full workload once: unmodified clust-mst --presketched on the same %d sketches, "
                              "generateMST phase (inverted-index compute_minhash_mst + Kruskal), index cached by a first run" % n,
                    "generate_mst_s": idx["runs"][0]["generate_mst_s"], "wall_s": idx["runs"][0]["wall_s"],
                    "first_run_generate_mst_s": idx["first_run"]["generate_mst_s"],
                    "minhash_distance": brute}
                # parity at full size: the reference's own edge.mst and cluster file against this run's MST
                ref = idx.get("result")
                if ref is not None:
                    ours_w = np.sort(mst["dist"])
                    lab = partition_labels(n, mst["pre"], mst["suf"], mst["dist"], thr)
                    rd = ref["dist"]
                    line["parity"] = {
                        "vs": "the unmodified reference on the same %d sketches in this run: cluster file of clust-mst "
                              "--presketched; MST weights of its compute_minhash_mst (called through oracle/_ref/ref_harness "
                              "msti, the binary itself does not write edge.mst in this mode)" % n,
                        "mst_edges": [int(len(mst)), None if rd is None else int(len(rd))],
                        "sorted_weights_bit_identical": None if rd is None else bool(
                            len(ours_w) == len(rd) and np.array_equal(ours_w.view(np.uint64), np.sort(rd).view(np.uint64))),
                        "partition_identical": bool(np.array_equal(lab, ref["labels"])),
                        "clusters": [int(len(np.unique(lab))), int(len(np.unique(ref["labels"])))]}
                    line["parity"]["ok"] = bool(line["parity"]["sorted_weights_bit_identical"] and line["parity"]["partition_identical"])
            else:
                line["cpu_baseline"] = brute
    if dist is not None:
        ctx.mg_leave()
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
    ap.add_argument("--no-robust", action="store_true", help="pairs workload: skip the robustness legs")
    ap.add_argument("--no-large", action="store_true", help="pairs workload: skip the large-set (scaling) leg")
    ap.add_argument("--dense-n", type=int, default=16384, help="robustness: sketches of the one-species set")
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
