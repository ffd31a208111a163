"""x as the REFERENCE's own GPU backend returns it, for the Nek matrices.

    gpurun -- python tests/golden/make_cusolver_golden.py     (needs a GPU)

The reference's `--solver cusolver` backend (src/cusparse.c:164-213: RCM
ordering :66-85, cusolverSpDcsrlsvchol :181-197, x scattered back through the
permutation :203-204) is the one backend of the reference that both builds in
this image and hands its solution back to the caller.  oracle/Makefile
(`make ref-cusolver`) compiles it from the reference's sources, untouched, into
oracle/_ref/libref_lsbench_cusolver.so; this script drives that library the way
lsbench_bench does (src/lsbench.c:156-187: x zeroed, r[i] = i) and stores x:

    gpurun_out/cusolver_x.npz  ->  copy to tests/golden/cusolver_x.npz

tests/test_oracle.py then holds the oracle's direct solve against these vectors
without a GPU, and tests/test_gpu_parity.py holds the b200 solve against a live
run of the same library.  Nothing of /root/reference is read at run time.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

REF_CU = os.path.join(orc.ROOT, "oracle", "_ref", "libref_lsbench_cusolver.so")


class _RefCsr(C.Structure):  # src/lsbench-impl.h:22-26
    _fields_ = [("nrows", C.c_uint), ("base", C.c_uint), ("offs", C.POINTER(C.c_uint)),
                ("cols", C.POINTER(C.c_uint)), ("vals", C.POINTER(C.c_double))]


_lib = None


def ref_lib():
    global _lib
    if _lib is None:
        R = C.CDLL(REF_CU)
        R.lsbench_matrix_read.restype = C.POINTER(_RefCsr)
        R.lsbench_matrix_read.argtypes = [C.c_char_p]
        R.lsbench_matrix_free.argtypes = [C.POINTER(_RefCsr)]
        R.lsbench_init.restype = C.c_void_p
        R.lsbench_init.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        R.lsbench_finalize.argtypes = [C.c_void_p]
        R.cusparse_bench.restype = C.c_int
        R.cusparse_bench.argtypes = [C.c_void_p, C.POINTER(_RefCsr), C.c_void_p, C.c_void_p]
        _lib = R
    return _lib


_cb = None


def reference_cusolver_solve(path):
    """x of the reference's cusparse_bench on the matrix file `path`, b[i] = i,
    default ordering (RCM, src/lsbench.c:95).  lsbench_init runs once per process
    (getopt state; it also runs cusparse_init, src/lsbench.c:143); the matrix name
    it was given only appears in the CSV row the backend prints."""
    global _cb
    R = ref_lib()
    if _cb is None:
        words = [b"driver", b"--solver", b"cusolver", b"--matrix", b"(several)", b"--trials=1"]
        argv = (C.c_char_p * (len(words) + 1))(*words, None)
        _cb = R.lsbench_init(len(words), argv)
        assert _cb, "reference lsbench_init failed"
    A = R.lsbench_matrix_read(path.encode())
    n = A.contents.nrows
    x, r = np.zeros(n), np.arange(n, dtype=np.float64)
    sys.stdout.flush()
    rc = R.cusparse_bench(x.ctypes.data, A, r.ctypes.data, _cb)
    R.lsbench_matrix_free(A)
    assert rc == 0, "reference cusparse_bench returned %d" % rc
    return x


def cusolver_rcm(path):
    """the permutation the reference's csr_init asks cuSOLVER for
    (src/cusparse.c:55-71: offsets carry the base, columns as in the file,
    descriptor base = the file's): the same call, made directly"""
    A = orc.matrix_read(path)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    so = C.CDLL(os.path.join(cuda, "lib64", "libcusolver.so"))
    sp = C.CDLL(os.path.join(cuda, "lib64", "libcusparse.so"))
    h, d = C.c_void_p(), C.c_void_p()
    assert so.cusolverSpCreate(C.byref(h)) == 0
    assert sp.cusparseCreateMatDescr(C.byref(d)) == 0
    assert sp.cusparseSetMatIndexBase(d, int(A.base)) == 0
    off = (A.offs.astype(np.int64) + A.base).astype(np.int32)
    col = A.cols.astype(np.int32)
    q = np.zeros(A.nrows, dtype=np.int32)
    so.cusolverSpXcsrsymrcmHost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
    rc = so.cusolverSpXcsrsymrcmHost(h, A.nrows, int(A.offs[-1]), d, off.ctypes.data, col.ctypes.data,
                                     q.ctypes.data)
    assert rc == 0, rc
    so.cusolverSpDestroy(h)
    return q


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(orc.ROOT, "gpurun_out", "cusolver_x.npz"))
    ap.add_argument("names", nargs="*", default=orc.NEK + ["I1_05x05"])
    a = ap.parse_args()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    out = {}
    for name in a.names:
        out[name] = reference_cusolver_solve(orc.matrix_path(name))
        print(name, out[name].shape, float(np.linalg.norm(out[name])), flush=True)
        np.savez_compressed(a.out, **out)
        try:
            out[name + "__rcm"] = cusolver_rcm(orc.matrix_path(name))
        except Exception as e:  # the permutation is extra evidence, x is the fixture
            print("no permutation for", name, ":", e, flush=True)
        np.savez_compressed(a.out, **out)   # after every matrix: a fatal errx() keeps the rest


if __name__ == "__main__":
    main()
