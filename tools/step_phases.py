"""Wall-clock phases of one multi-GPU bench step (pass, Boruvka phases, all-reduces); run under torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist
import rabbittclust_b200 as T
from rabbittclust_b200 import synth

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h, off = synth.family_sketches_fast(1000, 100, 1000, seed=7)
n = len(off) - 1
ctx = T.Context(lr)
ctx.upload(hashes=h, off=off)
L = T.lib()
radio = T.radio(0.05, 21)
best_key = torch.empty(n, dtype=torch.int64, device="cuda")
best_edge = torch.empty(n, dtype=torch.int64, device="cuda")
comp = np.arange(n, dtype=np.int32)
names = ["pass", "upload_comp", "phase1", "reduce1", "phase2", "reduce2"]
acc = {k: [] for k in names}
def red(t):
    if world > 1:
        t.bitwise_xor_(-2**63); dist.all_reduce(t, op=dist.ReduceOp.MIN); t.bitwise_xor_(-2**63)
    torch.cuda.current_stream().synchronize()
for it in range(25):
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    ts = [time.perf_counter()]
    ctx.allpairs_pass(radio, rank, world); ts.append(time.perf_counter())
    T._chk(L.tc_boruvka_upload_comp(ctx._h, comp.ctypes.data_as(C.POINTER(C.c_int32)), n)); ts.append(time.perf_counter())
    T._chk(L.tc_boruvka_phase1(ctx._h, 0, best_key.data_ptr())); ts.append(time.perf_counter())
    red(best_key); ts.append(time.perf_counter())
    T._chk(L.tc_boruvka_phase2(ctx._h, 0, best_key.data_ptr(), best_edge.data_ptr())); ts.append(time.perf_counter())
    red(best_edge); ts.append(time.perf_counter())
    if it >= 5:
        for k, a, b in zip(names, ts[:-1], ts[1:]):
            acc[k].append((b - a) * 1e3)
if rank == 0:
    print(" ".join("%s %.3f" % (k, float(np.median(v))) for k, v in acc.items()), "| kernel %.3f" % ctx.last_pass_kernel_ms)

# the same round queued back to back on the library's stream (tc_boruvka_set_async), as bench.py runs it
T._chk(L.tc_boruvka_set_async(ctx._h, 1))
stream = torch.cuda.ExternalStream(ctx.stream, device=lr)
def red_async(t):
    if world > 1:
        t.bitwise_xor_(-2**63); dist.all_reduce(t, op=dist.ReduceOp.MIN); t.bitwise_xor_(-2**63)
for variant in ("library stream", "library stream, no sign flip"):
    if variant != "library stream" and world > 1:
        try:                                   # does this torch/NCCL pair reduce uint64?
            dist.all_reduce(best_key.view(torch.uint64), op=dist.ReduceOp.MIN)
            torch.cuda.synchronize()
        except Exception as ex:
            if rank == 0:
                print("uint64 all_reduce unavailable:", str(ex)[:200])
            break
    tp, tr = [], []
    for it in range(25):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.allpairs_pass(radio, rank, world)
        t1 = time.perf_counter()
        with torch.cuda.stream(stream):
            T._chk(L.tc_boruvka_upload_comp(ctx._h, comp.ctypes.data_as(C.POINTER(C.c_int32)), n))
            T._chk(L.tc_boruvka_phase1(ctx._h, 0, best_key.data_ptr()))
            if variant == "library stream":
                red_async(best_key)
            elif world > 1:
                dist.all_reduce(best_key.view(torch.uint64), op=dist.ReduceOp.MIN)
            T._chk(L.tc_boruvka_phase2(ctx._h, 0, best_key.data_ptr(), best_edge.data_ptr()))
            if variant == "library stream":
                red_async(best_edge)
            elif world > 1:
                dist.all_reduce(best_edge.view(torch.uint64), op=dist.ReduceOp.MIN)
        stream.synchronize()
        t2 = time.perf_counter()
        if it >= 5:
            tp.append((t1 - t0) * 1e3); tr.append((t2 - t1) * 1e3)
    if rank == 0:
        print("%s: pass %.3f round %.3f" % (variant, float(np.median(tp)), float(np.median(tr))))
if world > 1:
    dist.destroy_process_group()
