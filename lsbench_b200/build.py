"""In-tree build of libb200.so (hand-written sm_100a kernels + the C ABI).

    python -m lsbench_b200.build [--force]

nvcc cross-compiles without a GPU; the .so stays next to this file so it
travels to the GPU box with the repository snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_obj")
LIB = os.path.join(PKG, "libb200.so")
SOURCES = ["ctx.cu", "convert.cu", "spmv.cu", "pcg.cu", "generate.cu",
           "dist.cu", "small.cu", "ingest.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx():
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); prefer the distro one
    for cand in (os.environ.get("HOSTCXX"), "/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def _digest(paths, extra):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    nvcc, cxx = _nvcc(), _host_cxx()
    os.makedirs(OBJ, exist_ok=True)
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    headers.append(os.path.join(ROOT, "include", "b200.h"))
    flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
             "-ccbin", cxx, "-I", os.path.join(ROOT, "include"), "-I", CSRC] + ARCH
    if verbose:
        flags += ["-Xptxas", "-v"]
    jobs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src + ".o")
        stamp = obj + ".sha"
        dig = _digest([path] + headers, " ".join(flags))
        fresh = (os.path.exists(obj) and os.path.exists(stamp)
                 and open(stamp).read() == dig)
        if force or not fresh:
            jobs.append((path, obj, stamp, dig))

    def compile_one(job):
        path, obj, stamp, dig = job
        r = subprocess.run([nvcc] + flags + ["-c", path, "-o", obj],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (path, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dig)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(compile_one, jobs))

    objs = [os.path.join(OBJ, s + ".o") for s in SOURCES]
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + \
              ["-ccbin", cxx, "-cudart", "static", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
