"""Quick wall-clock probe of the sketch kernel (development aid)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch, ctypes as C
import rabbittclust_b200 as T
ng, L = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
c = T.Context(0)
seq = torch.randint(0, 4, (ng * L + 64,), dtype=torch.uint8, device="cuda")
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
seq = lut[seq.long()] if ng * L < 2e8 else torch.cat([lut[seq[i:i + 10**8].long()] for i in range(0, len(seq), 10**8)])
rec_off = (np.arange(ng + 1, dtype=np.uint64) * np.uint64(L))
r2g = np.arange(ng, dtype=np.uint32); ss = np.full(ng, 1000, np.uint32)
out = torch.zeros((ng, 1000), dtype=torch.int64, device="cuda"); outn = torch.zeros(ng, dtype=torch.int32, device="cuda")
L_ = T.lib()
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    T._chk(L_.tc_sketch_minhash_dev(c._h, seq.data_ptr(), rec_off.ctypes.data_as(C.POINTER(C.c_uint64)),
           r2g.ctypes.data_as(C.POINTER(C.c_uint32)), ng, ng, 21, ss.ctypes.data_as(C.POINTER(C.c_uint32)), 42,
           out.data_ptr(), 1000, outn.data_ptr()))
    torch.cuda.synchronize(); dt = time.time() - t
    print("genomes %d x %d bp: %.3f s  %.1f genomes/s  %.2f Gbp/s" % (ng, L, dt, ng / dt, ng * L / dt / 1e9), flush=True)
print(outn[:4].tolist(), out[0, :3].tolist())
