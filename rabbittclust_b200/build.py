"""Build libtclust_cuda.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtclust_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr"]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "csrc", "*.cpp")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "csrc", "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "csrc", "build", os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src in sorted(glob.glob(os.path.join(HERE, "csrc", "*.cpp"))):     # host-only sources (AVX2 packer): plain g++
        obj = os.path.join(HERE, "csrc", "build", os.path.basename(src)[:-4] + ".o")
        cmd = [os.environ.get("CXX", "/usr/bin/g++"), "-O3", "-std=c++17", "-fPIC", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out = p.communicate()[0].decode()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on " + src)
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    # a shared library links with unresolved references: make an internal function that lost its definition a build error
    und = subprocess.run(["nm", "-D", "--undefined-only", LIB], stdout=subprocess.PIPE).stdout.decode()
    missing = [l.split()[-1] for l in und.splitlines() if "tc_" in l]
    if missing:
        os.remove(LIB)
        raise RuntimeError("libtclust_cuda.so has undefined internal symbols: " + ", ".join(missing))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
