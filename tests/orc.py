"""ctypes bindings for the CPU oracle (oracle/_build/liboracle.so) and, when it
was built, the reference's own reader (oracle/_ref/libref_lsbench.so).

Test infrastructure only: imported from tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs, never from lsbench_b200/.
"""
import ctypes as C
import lzma
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
DATA_DIR = os.path.join(ROOT, "tests", "data")
LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libref_lsbench.so")
REF_DRIVER = os.path.join(ORACLE_DIR, "_ref", "driver")

NEK = ["tj7a_A_12", "tj7a_A_15", "tj7a_A_18",
       "xn3b_A_10", "xn3b_A_12", "xn3b_A_15", "xn3b_A_18"]
TOY = ["A0_02x02", "A1_02x02", "I1_05x05"]


def build_oracle():
    """Compile the C restatement (and oracle/_ref when the reference tree is
    mounted).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref-cusolver"], check=True)
        if os.path.exists(os.path.join(ROOT, "lsbench_b200", "libb200.so")):
            # the drop-in for real: the reference tree + the five registration edits + b200.c
            subprocess.run([sys.executable, os.path.join(ORACLE_DIR, "dropin.py")], check=True,
                           stdout=subprocess.DEVNULL)


class _Csr(C.Structure):
    _fields_ = [("nrows", C.c_uint32), ("base", C.c_uint32),
                ("offs", C.POINTER(C.c_uint32)), ("cols", C.POINTER(C.c_uint32)),
                ("vals", C.POINTER(C.c_double))]


class _Op(C.Structure):
    _fields_ = [("n", C.c_uint64), ("offs", C.POINTER(C.c_uint64)),
                ("cols", C.POINTER(C.c_uint32)), ("vals", C.POINTER(C.c_double))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build_oracle()
        L = C.CDLL(LIB)
        L.orc_matrix_read.restype = C.POINTER(_Csr)
        L.orc_matrix_read.argtypes = [C.c_char_p]
        L.orc_matrix_free.argtypes = [C.POINTER(_Csr)]
        for f in ("orc_op_upper_mirror", "orc_op_full"):
            getattr(L, f).restype = C.POINTER(_Op)
            getattr(L, f).argtypes = [C.POINTER(_Csr)]
        L.orc_op_perm_lower_mirror.restype = C.POINTER(_Op)
        L.orc_op_perm_lower_mirror.argtypes = [C.POINTER(_Csr), C.POINTER(C.c_int32)]
        L.orc_op_free.argtypes = [C.POINTER(_Op)]
        L.orc_spmv.argtypes = [C.POINTER(_Op)] + [C.c_void_p] * 3
        L.orc_spmv_omp.argtypes = [C.POINTER(_Op)] + [C.c_void_p] * 2
        L.orc_spmv_fma.argtypes = [C.POINTER(_Op)] + [C.c_void_p] * 2
        for f in ("orc_pcg", "orc_pcg_omp"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.POINTER(_Op), C.c_void_p, C.c_void_p,
                                      C.c_double, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_pcg_refine32.restype = C.c_int
        L.orc_pcg_refine32.argtypes = [C.POINTER(_Op), C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                                       C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_double)]
        L.orc_true_relres.restype = C.c_double
        L.orc_true_relres.argtypes = [C.POINTER(_Op), C.c_void_p, C.c_void_p]
        L.orc_ldlt_factor.restype = C.c_void_p
        L.orc_ldlt_factor.argtypes = [C.POINTER(_Op), C.c_int]
        L.orc_ldlt_nnz.restype = C.c_uint64
        L.orc_ldlt_nnz.argtypes = [C.c_void_p]
        L.orc_ldlt_status.restype = C.c_int
        L.orc_ldlt_status.argtypes = [C.c_void_p]
        L.orc_ldlt_solve.argtypes = [C.c_void_p] * 3
        L.orc_ldlt_free.argtypes = [C.c_void_p]
        for f in ("orc_gen_poisson7", "orc_gen_poisson27"):
            getattr(L, f).restype = C.POINTER(_Op)
            getattr(L, f).argtypes = [C.c_uint32, C.c_uint64, C.c_uint64]
        L.orc_gen_powerlaw.restype = C.POINTER(_Op)
        L.orc_gen_powerlaw.argtypes = [C.c_uint64] * 4
        L.orc_powerlaw_rowlen.restype = C.c_uint32
        L.orc_powerlaw_rowlen.argtypes = [C.c_uint64] * 3
        L.orc_powerlaw_table.argtypes = [C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class HostCsr:
    """Numpy copy of `struct csr` (src/lsbench-impl.h:22-26)."""

    def __init__(self, nrows, base, offs, cols, vals):
        self.nrows, self.base = int(nrows), int(base)
        self.offs = np.ascontiguousarray(offs, dtype=np.uint32)
        self.cols = np.ascontiguousarray(cols, dtype=np.uint32)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)

    @property
    def nnz(self):
        return int(self.offs[-1])

    def as_struct(self):
        s = _Csr(self.nrows, self.base,
                 self.offs.ctypes.data_as(C.POINTER(C.c_uint32)),
                 self.cols.ctypes.data_as(C.POINTER(C.c_uint32)),
                 self.vals.ctypes.data_as(C.POINTER(C.c_double)))
        return s


class Op:
    """Numpy copy of an orc_op (0-based, 64-bit offsets)."""

    def __init__(self, n, offs, cols, vals, ncols=None):
        self.n = int(n)
        self.offs = np.ascontiguousarray(offs, dtype=np.uint64)
        self.cols = np.ascontiguousarray(cols, dtype=np.uint32)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.ncols = int(ncols) if ncols is not None else self.n

    @property
    def nnz(self):
        return int(self.offs[-1])

    def as_struct(self):
        return _Op(self.n, self.offs.ctypes.data_as(C.POINTER(C.c_uint64)),
                   self.cols.ctypes.data_as(C.POINTER(C.c_uint32)),
                   self.vals.ctypes.data_as(C.POINTER(C.c_double)))

    def scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals, self.cols.astype(np.int64),
                              self.offs.astype(np.int64)),
                             shape=(self.n, self.ncols))

    def rowlens(self):
        return np.diff(self.offs.astype(np.int64))


def _take_csr(p):
    a = p.contents
    nnz = int(a.offs[a.nrows])
    out = HostCsr(a.nrows, a.base, _np(a.offs, a.nrows + 1, np.uint32),
                  _np(a.cols, nnz, np.uint32), _np(a.vals, nnz, np.float64))
    return out


def _take_op(p, ncols=None):
    m = p.contents
    nnz = int(m.offs[m.n])
    out = Op(m.n, _np(m.offs, m.n + 1, np.uint64), _np(m.cols, nnz, np.uint32),
             _np(m.vals, nnz, np.float64), ncols)
    lib().orc_op_free(p)
    return out


def matrix_path(name, tmpdir=None):
    """Path of a COO text matrix; Nek files are stored xz-compressed and are
    unpacked once into tmpdir (default: $TMPDIR/lsbench_b200_data)."""
    plain = os.path.join(DATA_DIR, name + ".txt")
    if os.path.exists(plain):
        return plain
    packed = plain + ".xz"
    if not os.path.exists(packed):
        raise FileNotFoundError(name)
    import tempfile
    d = tmpdir or os.path.join(tempfile.gettempdir(), "lsbench_b200_data")
    os.makedirs(d, exist_ok=True)
    out = os.path.join(d, name + ".txt")
    if not os.path.exists(out):
        part = "%s.%d.part" % (out, os.getpid())  # ranks may unpack concurrently
        with lzma.open(packed, "rb") as f, open(part, "wb") as g:
            g.write(f.read())
        os.replace(part, out)
    return out


def matrix_read(path):
    p = lib().orc_matrix_read(path.encode())
    if not p:
        return None
    out = _take_csr(p)
    lib().orc_matrix_free(p)
    return out


def ref_matrix_read(path):
    """The reference's own lsbench_matrix_read (src/lsbench-csr.c:29), from
    oracle/_ref.  Returns None when _ref was not built."""
    if not os.path.exists(REF_LIB):
        return None
    R = C.CDLL(REF_LIB)
    R.lsbench_matrix_read.restype = C.POINTER(_Csr)
    R.lsbench_matrix_read.argtypes = [C.c_char_p]
    R.lsbench_matrix_free.argtypes = [C.POINTER(_Csr)]
    p = R.lsbench_matrix_read(path.encode())
    out = _take_csr(p)
    R.lsbench_matrix_free(p)
    return out


def op_upper_mirror(A):
    s = A.as_struct()
    return _take_op(lib().orc_op_upper_mirror(C.byref(s)))


def op_full(A):
    s = A.as_struct()
    return _take_op(lib().orc_op_full(C.byref(s)))


def op_perm_lower_mirror(A, q):
    """the operator the reference's cuSOLVER backend solves (oracle/operator.c)"""
    s = A.as_struct()
    q = np.ascontiguousarray(q, dtype=np.int32)
    return _take_op(lib().orc_op_perm_lower_mirror(C.byref(s), q.ctypes.data_as(C.POINTER(C.c_int32))))


def rhs(n):
    return np.arange(n, dtype=np.float64)  # src/lsbench.c:159-160


def spmv(M, x, want_abs=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(M.n)
    ya = np.empty(M.n) if want_abs else None
    s = M.as_struct()
    lib().orc_spmv(C.byref(s), x.ctypes.data, y.ctypes.data,
                   ya.ctypes.data if want_abs else None)
    return (y, ya) if want_abs else y


def spmv_fma(M, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(M.n)
    s = M.as_struct()
    lib().orc_spmv_fma(C.byref(s), x.ctypes.data, y.ctypes.data)
    return y


def pcg(M, b, x0=None, tol=1e-10, maxit=10000, omp=False):
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(M.n) if x0 is None else np.array(x0, dtype=np.float64)
    it, rel = C.c_int(0), C.c_double(0)
    s = M.as_struct()
    f = lib().orc_pcg_omp if omp else lib().orc_pcg
    rc = f(C.byref(s), b.ctypes.data, x.ctypes.data, tol, maxit,
           C.byref(it), C.byref(rel))
    return x, it.value, rel.value, rc


def cheb_lmax(M):
    s = M.as_struct()
    f = lib().orc_cheb_lmax
    f.restype = C.c_double
    return f(C.byref(s))


def pcg_cheb(M, b, degree=2, lmax=None, ratio=30.0, x0=None, tol=1e-10, maxit=10000):
    """Chebyshev-Jacobi preconditioned CG, oracle/krylov.c"""
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(M.n) if x0 is None else np.array(x0, dtype=np.float64)
    it, rel = C.c_int(0), C.c_double(0)
    s = M.as_struct()
    f = lib().orc_pcg_cheb
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                  C.POINTER(C.c_int), C.POINTER(C.c_double)]
    rc = f(C.byref(s), b.ctypes.data, x.ctypes.data, tol, maxit, degree,
           cheb_lmax(M) if lmax is None else lmax, ratio, C.byref(it), C.byref(rel))
    return x, it.value, rel.value, rc


def pcg_bj(M, b, block_of_row, x0=None, tol=1e-10, maxit=10000):
    """block-Jacobi preconditioned CG on a given partition of the rows, oracle/krylov.c"""
    b = np.ascontiguousarray(b, dtype=np.float64)
    blk = np.ascontiguousarray(block_of_row, dtype=np.uint32)
    assert blk.size == M.n
    x = np.zeros(M.n) if x0 is None else np.array(x0, dtype=np.float64)
    it, rel = C.c_int(0), C.c_double(0)
    s = M.as_struct()
    f = lib().orc_pcg_bj
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p,
                  C.POINTER(C.c_int), C.POINTER(C.c_double)]
    rc = f(C.byref(s), b.ctypes.data, x.ctypes.data, tol, maxit, blk.ctypes.data, C.byref(it), C.byref(rel))
    return x, it.value, rel.value, rc


def pcg_refine32(M, b, x0=None, tol=1e-10, maxit=10000, eta=1e-4):
    """fp32-stored operator + fp64 refinement, oracle/krylov.c"""
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(M.n) if x0 is None else np.array(x0, dtype=np.float64)
    it, outer, rel = C.c_int(0), C.c_int(0), C.c_double(0)
    s = M.as_struct()
    rc = lib().orc_pcg_refine32(C.byref(s), b.ctypes.data, x.ctypes.data, tol, maxit, eta,
                                C.byref(it), C.byref(outer), C.byref(rel))
    return x, it.value, outer.value, rel.value, rc


def true_relres(M, b, x):
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    s = M.as_struct()
    return lib().orc_true_relres(C.byref(s), b.ctypes.data, x.ctypes.data)


class Ldlt:
    def __init__(self, M, ordering=1):
        self._s = M.as_struct()
        self._M = M
        self.h = lib().orc_ldlt_factor(C.byref(self._s), ordering)
        self.n = M.n

    @property
    def nnz(self):
        return lib().orc_ldlt_nnz(self.h)

    @property
    def spd(self):
        return lib().orc_ldlt_status(self.h) == 0

    def solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.n)
        lib().orc_ldlt_solve(self.h, b.ctypes.data, x.ctypes.data)
        return x

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ldlt_free(self.h)
            self.h = None


def gen_poisson7(N, row0=0, row1=None):
    row1 = N ** 3 if row1 is None else row1
    return _take_op(lib().orc_gen_poisson7(N, row0, row1), ncols=N ** 3)


def gen_poisson27(N, row0=0, row1=None):
    row1 = N ** 3 if row1 is None else row1
    return _take_op(lib().orc_gen_poisson27(N, row0, row1), ncols=N ** 3)


def gen_powerlaw(n, seed, row0=0, row1=None):
    row1 = n if row1 is None else row1
    return _take_op(lib().orc_gen_powerlaw(n, seed, row0, row1), ncols=n)


def powerlaw_table():
    t = np.zeros(65537, dtype=np.uint64)
    lib().orc_powerlaw_table(t.ctypes.data)
    return t


def num_threads():
    return lib().orc_num_threads()


def set_threads(n=0):
    """OpenMP threads of the *_omp legs; 0 = every online core, whatever
    OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    return lib().orc_set_threads(int(n))


def pcg_slab_seconds(M, n_global, row0, its):
    """seconds for `its` Jacobi-PCG iterations' worth of passes over the row slab M
    (oracle/krylov.c orc_pcg_slab_seconds): bench.py's CPU timing sample"""
    s = M.as_struct()
    f = lib().orc_pcg_slab_seconds
    f.restype = C.c_double
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
    return f(C.byref(s), n_global, row0, its)
