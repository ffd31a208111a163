"""ctypes binding of include/b200.h (libb200.so).

This is plumbing for tests/ and bench.py; the product is the shared library.
There is no CPU fallback: if the library is missing, or no CUDA device is
usable, the calls raise.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(PKG, "libb200.so")   # (B200_LIB: A/B runs of two builds)

HIST_BINS = 24
NCCL_ID_BYTES = 128

MAT_SYM_UPPER = 1 << 0
MAT_FORCE_VECTOR = 1 << 1
MAT_FORCE_SELL = 1 << 2
MAT_NO_SORT = 1 << 3
MAT_NO_COMPRESS = 1 << 4
MAT_VALUES_F32 = 1 << 5
MAT_COL_BLOCK = 1 << 6

GEN_POISSON7, GEN_POISSON27, GEN_POWERLAW = 1, 2, 3

PCG_TIME_KERNELS = 1 << 0
PCG_NO_GRAPH = 1 << 1
PCG_NO_SMALL = 1 << 2
PCG_CHEBYSHEV2 = 1 << 4
PCG_CHEBYSHEV3 = 1 << 5
PCG_BLOCK_JACOBI = 1 << 6

# every symbol include/b200.h declares (tests check the library exports them)
SYMBOLS = [
    "b200_last_error", "b200_abi_version", "b200_device_count",
    "b200_ctx_create", "b200_nccl_unique_id", "b200_ctx_create_dist",
    "b200_ctx_destroy", "b200_ctx_set_stream", "b200_ctx_sync", "b200_ctx_rank",
    "b200_row_block",
    "b200_malloc", "b200_free", "b200_memcpy_h2d", "b200_memcpy_d2h",
    "b200_memset", "b200_host_alloc", "b200_host_free",
    "b200_mat_from_csr", "b200_mat_generate", "b200_mat_destroy",
    "b200_coo_to_csr", "b200_mat_from_coo", "b200_text_to_csr",
    "b200_mat_get_info", "b200_mat_export", "b200_mat_halo_cols",
    "b200_mat_inv_diag", "b200_mat_block_jacobi_partition", "b200_spmv", "b200_spmv_host", "b200_spmv_time",
    "b200_pcg_solve", "b200_pcg_solve_host", "b200_mat_algorithmic_bytes",
]


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200 error %d: %s" % (code, msg))
        self.code = code


class MatInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_global", "row_begin", "n_local", "n_halo", "nnz", "nnz_padded",
        "sell_rows", "sell_slices", "sell_sigma", "sell_max_width",
        "vec_rows", "vec_nnz", "long_rows", "long_nnz",
        "interior_begin", "interior_end")] + [
        ("hist", C.c_uint64 * HIST_BINS), ("max_row_len", C.c_uint64),
        ("pattern_symmetric", C.c_uint32), ("sell_perm", C.c_uint32),
        ("device_bytes", C.c_uint64), ("sell_uniform_slices", C.c_uint64),
        ("matrix_stream_bytes", C.c_uint64), ("values_f32", C.c_uint32), ("col_blocks", C.c_uint32)]


class PcgOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("maxit", C.c_int32),
                ("check_every", C.c_int32), ("flags", C.c_uint32)]


class PcgResult(C.Structure):
    _fields_ = [("iters", C.c_int32), ("status", C.c_int32),
                ("relres", C.c_double), ("true_relres", C.c_double),
                ("bnorm", C.c_double), ("solve_ms", C.c_float),
                ("spmv_ms", C.c_float), ("update_ms", C.c_float),
                ("pupdate_ms", C.c_float), ("kernel_launches", C.c_int32),
                ("path", C.c_int32), ("outer_iters", C.c_int32), ("replacements", C.c_int32),
                ("block_jacobi", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load():
    """dlopen libb200.so, building it first if only the sources are there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.b200_last_error.restype = C.c_char_p
    sig = {
        "b200_device_count": [C.POINTER(i32)],
        "b200_ctx_create": [i32, C.POINTER(vp)],
        "b200_nccl_unique_id": [vp],
        "b200_ctx_create_dist": [i32, i32, i32, vp, C.POINTER(vp)],
        "b200_ctx_destroy": [vp],
        "b200_ctx_set_stream": [vp, vp],
        "b200_ctx_sync": [vp],
        "b200_ctx_rank": [vp, C.POINTER(i32), C.POINTER(i32)],
        "b200_row_block": [u64, i32, i32, C.POINTER(u64), C.POINTER(u64)],
        "b200_malloc": [vp, C.c_size_t, C.POINTER(vp)],
        "b200_free": [vp, vp],
        "b200_memcpy_h2d": [vp, vp, vp, C.c_size_t],
        "b200_memcpy_d2h": [vp, vp, vp, C.c_size_t],
        "b200_memset": [vp, vp, i32, C.c_size_t],
        "b200_host_alloc": [C.c_size_t, C.POINTER(vp)],
        "b200_host_free": [vp],
        "b200_mat_from_csr": [vp, u32, u32, vp, vp, vp, u32, C.POINTER(vp)],
        "b200_mat_generate": [vp, i32, u64, u64, u32, C.POINTER(vp)],
        "b200_coo_to_csr": [vp, u64, vp, vp, vp, C.POINTER(u32), C.POINTER(u64), vp, vp, vp],
        "b200_mat_from_coo": [vp, u64, u32, vp, vp, vp, u32, C.POINTER(vp)],
        "b200_text_to_csr": [vp, C.c_char_p, u64, u64, C.POINTER(u32), C.POINTER(u64), vp, vp, vp,
                             C.POINTER(u64)],
        "b200_mat_destroy": [vp],
        "b200_mat_get_info": [vp, C.POINTER(MatInfo)],
        "b200_mat_export": [vp, vp, vp, vp],
        "b200_mat_halo_cols": [vp, vp],
        "b200_mat_inv_diag": [vp, vp],
        "b200_mat_block_jacobi_partition": [vp, vp, vp],
        "b200_spmv": [vp, vp, vp],
        "b200_spmv_host": [vp, vp, vp],
        "b200_spmv_time": [vp, vp, vp, i32, C.POINTER(C.c_float)],
        "b200_pcg_solve": [vp, vp, vp, C.POINTER(PcgOpts), C.POINTER(PcgResult)],
        "b200_pcg_solve_host": [vp, vp, vp, C.POINTER(PcgOpts), C.POINTER(PcgResult)],
        "b200_mat_algorithmic_bytes": [vp, C.POINTER(u64), C.POINTER(u64)],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes, f.restype = args, C.c_int
    L.b200_abi_version.restype = C.c_int
    _lib = L
    return L


def _chk(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise B200Error(rc, load().b200_last_error().decode(errors="replace"))
    return rc


def device_count():
    n = C.c_int(0)
    _chk(load().b200_device_count(C.byref(n)))
    return n.value


def row_block(n, rank, nranks):
    """[r0, r1) owned by `rank` (host arithmetic, no GPU)."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    _chk(load().b200_row_block(n, rank, nranks, C.byref(a), C.byref(b)))
    return a.value, b.value


def nccl_unique_id():
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _chk(load().b200_nccl_unique_id(buf))
    return bytes(buf.raw)


def coo_to_csr(ctx, rows, cols, vals):
    """Device-side restatement of the reader body (src/lsbench-csr.c:54-86):
    returns (nrows, offs u32, cols u32 with the base kept, vals)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    nnz = rows.size
    offs = np.empty(nnz + 1, dtype=np.uint32)
    oc = np.empty(max(nnz, 1), dtype=np.uint32)
    ov = np.empty(max(nnz, 1), dtype=np.float64)
    nr, m = C.c_uint32(0), C.c_uint64(0)
    _chk(load().b200_coo_to_csr(ctx.h, nnz, rows.ctypes.data, cols.ctypes.data,
                                vals.ctypes.data, C.byref(nr), C.byref(m),
                                offs.ctypes.data, oc.ctypes.data, ov.ctypes.data))
    return nr.value, offs[:nr.value + 1].copy(), oc[:m.value].copy(), ov[:m.value].copy()


def text_to_csr(ctx, body, nnz):
    """COO text (everything after the header line, bytes) -> (nrows, offs, cols,
    vals, lines parsed by the host); lines and numbers are parsed on the device."""
    offs = np.empty(nnz + 1, dtype=np.uint32)
    oc = np.empty(max(nnz, 1), dtype=np.uint32)
    ov = np.empty(max(nnz, 1), dtype=np.float64)
    nr, m, nh = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    _chk(load().b200_text_to_csr(ctx.h, body, len(body), nnz, C.byref(nr), C.byref(m),
                                 offs.ctypes.data, oc.ctypes.data, ov.ctypes.data, C.byref(nh)))
    return nr.value, offs[:nr.value + 1].copy(), oc[:m.value].copy(), ov[:m.value].copy(), nh.value


class DeviceArray:
    """n doubles of library-owned device memory."""

    def __init__(self, ctx, n):
        self.ctx, self.n = ctx, int(n)
        p = C.c_void_p()
        _chk(load().b200_malloc(ctx.h, self.n * 8, C.byref(p)))
        self.ptr = p.value

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.n
        _chk(load().b200_memcpy_h2d(self.ctx.h, self.ptr, a.ctypes.data, self.n * 8))
        return self

    def download(self):
        out = np.empty(self.n, dtype=np.float64)
        _chk(load().b200_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, self.n * 8))
        return out

    def zero(self):
        _chk(load().b200_memset(self.ctx.h, self.ptr, 0, self.n * 8))
        return self

    def free(self):
        if self.ptr:
            load().b200_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None):
        h = C.c_void_p()
        if nranks == 1:
            _chk(load().b200_ctx_create(device, C.byref(h)))
        else:
            _chk(load().b200_ctx_create_dist(device, rank, nranks, nccl_id, C.byref(h)))
        self.h, self.rank, self.nranks, self.device = h, rank, nranks, device

    def set_stream(self, cuda_stream):
        _chk(load().b200_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        _chk(load().b200_ctx_sync(self.h))

    def array(self, n):
        return DeviceArray(self, n)

    def close(self):
        if self.h:
            load().b200_ctx_destroy(self.h)
            self.h = None


class Matrix:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    @classmethod
    def from_csr(cls, ctx, nrows, base, offs, cols, vals, flags=0):
        """Host CSR with the field meanings of `struct csr`
        (src/lsbench-impl.h:22-26): offs 0-based, cols carrying `base`."""
        offs = np.ascontiguousarray(offs, dtype=np.uint32)
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        h = C.c_void_p()
        _chk(load().b200_mat_from_csr(ctx.h, nrows, base, offs.ctypes.data,
                                      cols.ctypes.data, vals.ctypes.data, flags,
                                      C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_coo(cls, ctx, base, rows, cols, vals, flags=0):
        """COO records in file order (src/lsbench-csr.c:49-53) -> device
        layout; sort / fold / row-compress on the device."""
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        h = C.c_void_p()
        _chk(load().b200_mat_from_coo(ctx.h, rows.size, base, rows.ctypes.data,
                                      cols.ctypes.data, vals.ctypes.data, flags,
                                      C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def generate(cls, ctx, kind, size, seed=0, flags=0):
        h = C.c_void_p()
        _chk(load().b200_mat_generate(ctx.h, kind, size, seed, flags, C.byref(h)))
        return cls(ctx, h)

    def info(self):
        i = MatInfo()
        _chk(load().b200_mat_get_info(self.h, C.byref(i)))
        return i

    def algorithmic_bytes(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _chk(load().b200_mat_algorithmic_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def export(self):
        i = self.info()
        offs = np.empty(i.n_local + 1, dtype=np.uint64)
        cols = np.empty(max(i.nnz, 1), dtype=np.uint32)
        vals = np.empty(max(i.nnz, 1), dtype=np.float64)
        _chk(load().b200_mat_export(self.h, offs.ctypes.data, cols.ctypes.data,
                                    vals.ctypes.data))
        return offs, cols[:i.nnz], vals[:i.nnz]

    def halo_cols(self):
        i = self.info()
        g = np.empty(max(i.n_halo, 1), dtype=np.uint64)
        _chk(load().b200_mat_halo_cols(self.h, g.ctypes.data))
        return g[:i.n_halo]

    def block_jacobi_partition(self):
        """(block id of every row, block size) of B200_PCG_BLOCK_JACOBI; size 0 = not in use"""
        blk = np.zeros(max(self.info().n_local, 1), dtype=np.uint32)
        bs = C.c_uint32(0)
        _chk(load().b200_mat_block_jacobi_partition(self.h, blk.ctypes.data, C.byref(bs)))
        return blk[:self.info().n_local], bs.value

    def inv_diag(self):
        d = np.empty(self.info().n_local, dtype=np.float64)
        _chk(load().b200_mat_inv_diag(self.h, d.ctypes.data))
        return d

    def spmv_host(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.info().n_local, dtype=np.float64)
        _chk(load().b200_spmv_host(self.h, x.ctypes.data, y.ctypes.data))
        return y

    def spmv(self, dx, dy):
        _chk(load().b200_spmv(self.h, _ptr(dx), _ptr(dy)))

    def spmv_time(self, dx, dy, reps=20):
        ms = C.c_float(0)
        _chk(load().b200_spmv_time(self.h, _ptr(dx), _ptr(dy), reps, C.byref(ms)))
        return ms.value

    def pcg_host(self, b, x0=None, tol=1e-10, maxit=10000, check_every=0, flags=0):
        """The X_bench call shape: host r in, host x out."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
        o, r = PcgOpts(tol, maxit, check_every, flags), PcgResult()
        rc = _chk(load().b200_pcg_solve_host(self.h, b.ctypes.data, x.ctypes.data,
                                             C.byref(o), C.byref(r)), allow=(5,))
        return x, r, rc

    def pcg(self, db, dx, tol=1e-10, maxit=10000, check_every=0, flags=0):
        o, r = PcgOpts(tol, maxit, check_every, flags), PcgResult()
        rc = _chk(load().b200_pcg_solve(self.h, _ptr(db), _ptr(dx), C.byref(o),
                                        C.byref(r)), allow=(5,))
        return r, rc

    def close(self):
        if self.h:
            load().b200_mat_destroy(self.h)
            self.h = None


def _ptr(a):
    if isinstance(a, DeviceArray):
        return C.c_void_p(a.ptr)
    if hasattr(a, "data_ptr"):  # a torch CUDA tensor
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))
