"""The product's Jacobi-PCG kernels (lsbench_b200/csrc/pcg_kernels.cuh) and its
SELL SpMV with the fused dot product, compiled for the host and run on the SIMT
emulator of tests/simt_emul.hpp -- one fiber per CUDA thread, barriers for
__syncthreads and the warp shuffles, the grid-wide fixed-order reductions
included.  The solve they produce is held against the oracle: same iteration
count, same solution, the stopping rules, bit-reproducibility, for the
three-kernel iteration.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc
from test_spmv_emul import sellc_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libpcg_emul.so")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("no CUDA headers")
    subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-w", "-DB2_SIMT_EMUL", "-shared", "-fPIC",
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "lsbench_b200", "csrc"),
                    "-I", os.path.join(cuda, "include"), "-I", os.path.join(ROOT, "tests"),
                    os.path.join(ROOT, "tests", "pcg_emul.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.emul_pcg.restype = C.c_int
    L.emul_pcg.argtypes = ([C.c_uint32, C.c_uint32] + [C.c_void_p] * 8 + [C.c_double, C.c_int, C.c_int,
                           C.c_uint, C.c_uint] + [C.POINTER(C.c_int)] * 2 + [C.POINTER(C.c_double)]
                           + [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)])
    return L


LAST = {}   # replacements / true_relres of the most recent solve()


def solve(emul, M, b, x0=None, tol=1e-10, maxit=10000, sr=False, kernel=0):
    Lay = sellc_layout(M)
    vals = Lay["vals"].astype(np.float64)
    assert np.array_equal(vals[:-1].astype(np.float32), Lay["vals"][:-1])
    S = M.scipy()
    d = S.diagonal()
    dinv = np.where(d != 0, 1.0 / np.where(d != 0, d, 1.0), 1.0)
    x = np.zeros(M.n) if x0 is None else np.array(x0, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    it, st, rel = C.c_int(0), C.c_int(0), C.c_double(0)
    rep, trr = C.c_int(0), C.c_double(0)
    p = lambda a: None if a is None else a.ctypes.data
    grid_spmv = (Lay["ns"] + 7) // 8
    grid_ew = (M.n + 255) // 256
    assert emul.emul_pcg(M.n, Lay["ns"], p(Lay["meta"]), p(Lay["ecols"]), p(Lay["dcols"]), p(vals), None,
                         p(dinv), p(b), p(x), tol, maxit, int(sr), grid_spmv, grid_ew,
                         C.byref(it), C.byref(st), C.byref(rel), p(Lay["vals"]), kernel,
                         3,
                         C.byref(rep), C.byref(trr)) == 0
    LAST["replacements"], LAST["true_relres"] = rep.value, trr.value
    return x, it.value, st.value, rel.value


@pytest.mark.parametrize("sr", [False])
def test_product_pcg_kernels_on_the_emulator(emul, sr):
    M = orc.gen_poisson7(12)                      # 1728 rows: 7 CTAs per kernel
    b = orc.rhs(M.n)
    x, it, st, rel = solve(emul, M, b, sr=sr)
    xo, ito, relo, rco = orc.pcg(M, b)
    assert st == 0 and rco == 0 and abs(it - ito) <= 1 and rel <= 1e-10
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-10
    assert orc.true_relres(M, b, x) <= 1e-10
    # bit-reproducible: every reduction has a fixed order
    x2, it2, _, _ = solve(emul, M, b, sr=sr)
    assert it2 == it and x2.tobytes() == x.tobytes()


@pytest.mark.parametrize("sr", [False])
def test_product_pcg_stopping_rules_on_the_emulator(emul, sr):
    M = orc.gen_poisson27(6)                      # 216 rows, one CTA
    b = orc.rhs(M.n)
    xs, its, st, _ = solve(emul, M, b, sr=sr)
    assert st == 0
    x, it, st, _ = solve(emul, M, b, maxit=5, sr=sr)          # stops at maxit, says so
    xo, ito, _, rco = orc.pcg(M, b, maxit=5)
    assert (it, st) == (5, 1) == (ito, rco) and np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-12
    x, it, st, _ = solve(emul, M, b, x0=xs, tol=1e-9, sr=sr)  # starting at the solution
    assert (it, st) == (0, 0)
    x, it, st, _ = solve(emul, M, b, maxit=8, sr=sr)          # maxit on a chunk boundary
    assert (it, st) == (8, 1)
    x, it, st, _ = solve(emul, M, b, maxit=its, sr=sr)        # converges exactly at maxit
    assert (it, st) == (its, 0) and x.tobytes() == xs.tobytes()


@pytest.mark.parametrize("gen,N", [("poisson27", 8), ("poisson7", 10)])
def test_every_spmv_kernel_inside_the_solve(emul, gen, N):
    """the same solve with the SpMV + fused dot on fp64 and on fp32-stored values: fp32
    storage is exact for the stencils and changes neither the row sums nor the order
    of the dot product, so the iterates are bit-identical"""
    M = getattr(orc, "gen_" + gen)(N)
    b = orc.rhs(M.n)
    runs = [solve(emul, M, b, kernel=k) for k in range(2)]
    assert all(r[2] == 0 for r in runs)
    assert runs[0][0].tobytes() == runs[1][0].tobytes() and runs[0][1] == runs[1][1]
    assert orc.true_relres(M, b, runs[0][0]) <= 1e-10


def nek_solve(emul, name, sr=False, tol=1e-10, maxit=5000):
    """A Nek coarse-grid operator (CHOLMOD's: upper triangle mirrored) through the
    product's streaming kernels on the emulator.  The fp64 Nek values are kept as
    they are (the layout helper's fp32 copy is not used here).
    -> M, b, x, iterations, status, recurrence relres, replacements, true relres"""
    A = orc.matrix_read(orc.matrix_path(name))
    M = orc.op_upper_mirror(A)
    b = orc.rhs(M.n)
    Lay = sellc_layout(M)
    # sellc_layout rounds to fp32 for the fp32-value kernels; rebuild the fp64 stream
    vals = np.zeros(Lay["vals"].size)
    o = 0
    for s_ in range(Lay["ns"]):
        w = int(Lay["meta"][s_, 1] & 0x7FFFFFFF)
        for l in range(32):
            r = 32 * s_ + l
            if r < M.n:
                a, e = int(M.offs[r]), int(M.offs[r + 1])
                vals[32 * o + l:32 * (o + e - a) + l:32] = M.vals[a:e]
        o += w
    d = M.scipy().diagonal()
    dinv = 1.0 / d
    x = np.zeros(M.n)
    it, st, rel = C.c_int(0), C.c_int(0), C.c_double(0)
    rep, trr = C.c_int(0), C.c_double(0)
    p = lambda a: None if a is None else a.ctypes.data
    assert emul.emul_pcg(M.n, Lay["ns"], p(Lay["meta"]), p(Lay["ecols"]), p(Lay["dcols"]), p(vals), None,
                         p(dinv), p(b), p(x), tol, maxit, int(sr), (Lay["ns"] + 7) // 8, (M.n + 255) // 256,
                         C.byref(it), C.byref(st), C.byref(rel), None, 0, 32,
                         C.byref(rep), C.byref(trr)) == 0
    return M, b, x, it.value, st.value, rel.value, rep.value, trr.value


@pytest.mark.parametrize("name,sr", [("tj7a_A_18", False), ("xn3b_A_18", False), ("xn3b_A_10", False),
                                     ("tj7a_A_12", False)])
def test_product_kernels_solve_a_nek_matrix_on_the_emulator(emul, name, sr):
    """BASELINE.json config 2 without a GPU: the product's streaming kernels (SELL
    SpMV + fused dot, K2, K3) on a Nek coarse-grid operator, ~15
    CTAs per kernel and ~300 iterations, against the SuperLU direct solve (the
    1e-8 parity bar) and the oracle's iteration count."""
    M, b, x, it, st, rel, rep, trr = nek_solve(emul, name, sr)
    _, ito, _, _ = orc.pcg(M, b)
    assert st == 0 and abs(it - ito) <= 3 and rel <= 1e-10
    if not sr:
        # same kernels, same launch geometry (15 - 26 CTAs: fewer than the GPU holds at
        # once), same fixed-order reductions: the emulator takes exactly the iterations
        # the B200 took (tests/golden/gpu_iters.json, from profiles/r01_nek_table...)
        import json
        gpu = json.load(open(os.path.join(ROOT, "tests", "golden", "gpu_iters.json")))["stream_iters"]
        assert it == gpu[name] and rep == 0
    assert orc.true_relres(M, b, x) <= 1e-10
    g = np.load(os.path.join(ROOT, "tests", "golden", "direct.npz"))[name]
    assert np.linalg.norm(x - g) / np.linalg.norm(g) <= 1e-8


def test_residual_replacement_on_the_emulator(emul):
    """The stopping test watches the recurrence residual; the bar is on b - A x.  At a
    tolerance where the two have drifted apart (2e-12 on tj7a_A_18: cond 2.5e4) the
    exit check finds ||b - A x|| above the bar, r is replaced by the true residual,
    p keeps its direction, and a handful of iterations later the bar is met by BOTH.
    Below what fp64 reaches for this system (3e-13) the tail after a replacement is
    bounded and the solve says so: status 4, not 5000 iterations."""
    M, b, x, it, st, rel, rep, trr = nek_solve(emul, "tj7a_A_18", False, 2e-12)
    assert st == 0 and rep >= 1 and rel <= 2e-12 and trr <= 2e-12
    assert orc.true_relres(M, b, x) <= 2e-12
    xo, ito, relo, rco = orc.pcg(M, b, tol=2e-12)        # the oracle states the same rule
    assert rco == 0 and orc.true_relres(M, b, xo) <= 2e-12 and abs(it - ito) <= 8
    M, b, x, it2, st2, rel2, rep2, trr2 = nek_solve(emul, "tj7a_A_18", False, 3e-13)
    assert st2 == 4 and rep2 >= 1 and it2 < it + 120 and trr2 > 3e-13
    _, ito2, _, rco2 = orc.pcg(M, b, tol=3e-13)
    assert rco2 == 4 and ito2 < ito + 120


@pytest.mark.parametrize("long_kernel", [0, 1])
def test_row_major_kernels_on_the_emulator(emul, long_kernel):
    """the bins of the power-law tail (BASELINE.json config 5): k_spmv_vec (one
    warp per row, 128-bit loads, butterfly reduction) and k_spmv_long (one CTA per
    row) on rows of 257 ... 1500 entries padded to multiples of 4 -- within the
    1e-13 bar of the CSR product (another summation tree, so not bit for bit), the
    fused dot product included"""
    emul.emul_rowmajor.argtypes = [C.c_int, C.c_uint, C.c_uint32] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
    rng = np.random.default_rng(23)
    n, nrows = 6000, 37
    lens = rng.integers(257, 1500, nrows)
    ids = rng.choice(n, nrows, replace=False).astype(np.uint32)
    pad = (lens + 3) // 4 * 4
    off = np.concatenate([[0], np.cumsum(pad)]).astype(np.uint64)
    o = off.astype(np.int64)
    lens = lens.astype(np.int64)
    cols = np.zeros(int(off[-1]) + 4, dtype=np.uint32)
    vals = np.zeros(int(off[-1]) + 4)
    for r in range(nrows):
        c = np.sort(rng.choice(n, lens[r], replace=False))
        cols[o[r]:o[r] + lens[r]] = c
        cols[o[r] + lens[r]:o[r + 1]] = ids[r]            # padding: own row, value 0
        vals[o[r]:o[r] + lens[r]] = rng.uniform(-1, 1, lens[r])
    x = rng.standard_normal(n)
    want = np.array([vals[o[r]:o[r + 1]] @ x[cols[o[r]:o[r + 1]]] for r in range(nrows)])
    scale = np.array([np.abs(vals[o[r]:o[r + 1]] * x[cols[o[r]:o[r + 1]]]).sum() for r in range(nrows)])
    p = lambda a: a.ctypes.data
    for dot in (0, 1):
        for grid in (1, 5):
            y = np.full(n, np.nan)
            d = np.zeros(1)
            assert emul.emul_rowmajor(long_kernel, grid, nrows, p(ids), p(off), p(cols), p(vals), p(x), p(y),
                                      dot, p(d)) == 0
            assert np.all(np.abs(y[ids] - want) <= 1e-13 * scale)
            assert np.isnan(np.delete(y, ids)).all()          # nothing else is written
            if dot:
                assert abs(d[0] - y[ids] @ x[ids]) <= 1e-12 * np.abs(y[ids] * x[ids]).sum()


@pytest.mark.parametrize("sr", [False])
def test_breakdown_is_reported_by_the_product_kernels(emul, sr):
    """[[1, 2], [2, 1]] x = [1, -1]: p.Ap = -2 in the first iteration -- the
    guard of K2 (pq <= 0) / K2' (delta - beta gamma / alpha_prev <= 0) sets status
    2 and leaves x alone; NaN in b does the same instead of spinning to maxit"""
    M = orc.Op(2, np.array([0, 2, 4], dtype=np.uint64), np.array([0, 1, 0, 1], dtype=np.uint32),
               np.array([1.0, 2.0, 2.0, 1.0]))
    x, it, st, _ = solve(emul, M, np.array([1.0, -1.0]), sr=sr)
    assert (it, st) == (0, 2) and x.tolist() == [0.0, 0.0]
    M = orc.gen_poisson7(4)
    b = orc.rhs(M.n)
    b[5] = np.nan
    x, it, st, _ = solve(emul, M, b, sr=sr, maxit=50)
    assert st == 2 and it <= 1


@pytest.mark.parametrize("gen,N", [("poisson27", 10), ("poisson7", 12)])
def test_two_launch_spmv_with_one_fused_dot(emul, gen, N):
    """what the overlapped multi-GPU SpMV launches: interior slices while the halo
    is in flight, boundary slices after it, ONE p.Ap -- the partial slots and the
    ticket span both launches: y with the bits of the CSR product, x.y to rounding,
    the ticket back at zero"""
    emul.emul_spmv_two_phase.argtypes = ([C.c_uint32, C.c_uint32] + [C.c_void_p] * 5 + [C.c_int, C.c_int,
                                         C.c_uint32, C.c_uint32] + [C.c_void_p] * 3)
    M = getattr(orc, "gen_" + gen)(N)
    Lay = sellc_layout(M)
    vals = Lay["vals"].astype(np.float64)
    x = np.random.default_rng(9).standard_normal(M.n)
    want = orc.spmv_fma(M, x)
    ns = Lay["ns"]
    wmax = 0
    p = lambda a: a.ctypes.data
    for kernel in (0,):
        for ib, ie in ((ns // 4, 3 * ns // 4), (0, ns - 1), (1, ns)):
            y, d = np.full(M.n, np.nan), np.zeros(1)
            assert emul.emul_spmv_two_phase(M.n, ns, p(Lay["meta"]), p(Lay["ecols"]), p(Lay["dcols"]), p(vals),
                                            p(Lay["vals"]), kernel, wmax, ib, ie, p(x), p(y), p(d)) == 0
            assert y.tobytes() == want.tobytes()
            assert abs(d[0] - x @ want) <= 1e-12 * np.abs(x * want).sum()


def test_column_blocked_spmv_on_the_emulator(emul):
    """B200_MAT_COL_BLOCK without a GPU: the product's kernels cut a power-law
    operator (oracle generator, BASELINE.json config 5) into column ranges
    (k_colblock_count / k_colblock_fill), every range becomes a SELL layout over all
    rows, and y = A_0 x; y += A_1 x; ... runs through the accumulate instantiation
    of the SELL kernel.  The pieces add up to the operator bit for bit; the product
    is the CSR product to the 1e-13 bar (row sums are formed block by block)."""
    emul.emul_colblock_split.argtypes = ([C.c_uint64] + [C.c_void_p] * 4 + [C.c_uint32, C.c_uint64, C.c_uint64]
                                         + [C.c_void_p] * 4 + [C.c_int, C.c_void_p])
    emul.emul_sell_acc.argtypes = [C.c_int, C.c_uint, C.c_uint32] + [C.c_void_p] * 5 + [C.c_uint32]
    n = 3000
    M = orc.gen_powerlaw(n, 2)
    keep = M.rowlens() <= 200                      # the SELL bin (longer rows go to other kernels)
    offs = M.offs.astype(np.int64)
    sel = np.concatenate([np.arange(offs[i], offs[i + 1]) if keep[i] else [] for i in range(n)]).astype(np.int64)
    lens = np.where(keep, M.rowlens(), 0)
    M = orc.Op(n, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64), M.cols[sel], M.vals[sel])
    width, nb = 704, 5                             # 5 ranges of 704 columns (a multiple of 32)
    assert (nb - 1) * width < n <= nb * width
    p = lambda a: a.ctypes.data
    cuts = np.minimum(np.arange(nb + 1, dtype=np.uint64) * np.uint64(width), np.uint64(n))
    cnt = np.zeros(nb * (n + 1), dtype=np.uint64)
    bad = np.zeros(1, dtype=np.uint32)
    emul.emul_colblock_split(n, p(M.offs), p(M.cols), p(M.vals), p(cuts), nb, n, 0, p(cnt), None, None, None, 0,
                             p(bad))
    assert bad[0] == 0
    cnt = cnt.reshape(nb, n + 1)
    want_cnt = np.array([[np.sum((M.cols[offs2[0]:offs2[1]] // width) == b) for offs2 in
                          zip(M.offs[:-1].astype(np.int64), M.offs[1:].astype(np.int64))] for b in range(nb)])
    assert np.array_equal(cnt[:, :n], want_cnt) and not cnt[:, n].any()
    boffs = np.concatenate([np.zeros((nb, 1), dtype=np.uint64), np.cumsum(cnt[:, :n], axis=1).astype(np.uint64)],
                           axis=1)             # the exclusive scan the host code does with cub
    boffs = np.ascontiguousarray(boffs)
    ocols = [np.zeros(max(int(boffs[b, n]), 1), dtype=np.uint32) for b in range(nb)]
    ovals = [np.zeros(max(int(boffs[b, n]), 1)) for b in range(nb)]
    pc = (C.c_void_p * nb)(*[p(a) for a in ocols])
    pv = (C.c_void_p * nb)(*[p(a) for a in ovals])
    emul.emul_colblock_split(n, p(M.offs), p(M.cols), p(M.vals), p(cuts), nb, n, 0, None, p(boffs), pc, pv, 1, None)
    # the pieces, row by row and in range order, are the operator
    S = M.scipy()
    x = np.random.default_rng(6).standard_normal(n)
    y = np.full(n, np.nan)
    yg = np.full(n, np.nan)                        # the same through the grouped kernel
    total = 0
    for b in range(nb):
        nnz_b = int(boffs[b, n])
        Mb = orc.Op(n, boffs[b], ocols[b][:max(nnz_b, 1)], ovals[b][:max(nnz_b, 1)], ncols=n)
        if nnz_b:
            assert Mb.cols[:nnz_b].min() >= b * width and Mb.cols[:nnz_b].max() < (b + 1) * width
        total += nnz_b
        Lay = sellc_layout(Mb)
        vals = np.zeros(Lay["allcols"].size)
        o = 0
        for s_ in range(Lay["ns"]):
            w = int(Lay["sell_off"][s_ + 1] - Lay["sell_off"][s_])
            for l in range(32):
                r = 32 * s_ + l
                if r < n:
                    a, e = int(Mb.offs[r]), int(Mb.offs[r + 1])
                    vals[32 * o + l:32 * (o + e - a) + l:32] = Mb.vals[a:e]
            o += w
        emul.emul_sell_acc(int(b > 0), 3, Lay["ns"], p(Lay["sell_off"]), p(Lay["allcols"]), p(vals), p(x), p(y), n)
        emul.emul_sell_acc(2 + int(b > 0), 2, Lay["ns"], p(Lay["sell_off"]), p(Lay["allcols"]), p(vals), p(x), p(yg), n)
        # four slices per warp trip: every row still adds its entries left to right
        assert yg.tobytes() == y.tobytes(), b
    assert total == M.nnz
    ref, scale = orc.spmv(M, x, want_abs=True)
    assert np.all(np.abs(y - ref) <= 1e-13 * np.maximum(scale, 1e-300))
    # an unsorted row is noticed
    cols2 = M.cols.copy()
    i = int(np.argmax(M.rowlens() >= 2))
    a = int(M.offs[i])
    cols2[a], cols2[a + 1] = cols2[a + 1], cols2[a]
    bad[:] = 0
    scratch = np.zeros(nb * (n + 1), dtype=np.uint64)
    emul.emul_colblock_split(n, p(M.offs), p(cols2), p(M.vals), p(cuts), nb, n, 0, p(scratch), None, None, None, 0,
                             p(bad))
    assert bad[0] == 1
    # ---- the same cut on a rank's renumbered columns (dist.cu: owned -> [0, n_own), remote ->
    # n_own + slot, slots in ascending global order, the first n_low of them below the own rows):
    # ranges follow the GLOBAL order [slots below | owned | slots above] and never cross a seam
    r0, r1 = 1024, 2048                            # this rank's rows
    n_own = r1 - r0
    rows = slice(int(M.offs[r0]), int(M.offs[r1]))
    gcols = M.cols[rows].astype(np.int64)
    remote = np.unique(gcols[(gcols < r0) | (gcols >= r1)])
    n_low = int(np.sum(remote < r0))
    slot = {int(g): i for i, g in enumerate(remote)}
    lcols = np.array([g - r0 if r0 <= g < r1 else n_own + slot[int(g)] for g in gcols], dtype=np.uint32)
    loffs = (M.offs[r0:r1 + 1] - M.offs[r0]).astype(np.uint64)
    ncols = n_own + len(remote)
    seams = [0, n_low // 2, n_low, n_low + n_own // 2, n_low + n_own, ncols]
    cuts2 = np.array(sorted(set(seams)), dtype=np.uint64)
    nb2 = len(cuts2) - 1
    cnt2 = np.zeros(nb2 * (n_own + 1), dtype=np.uint64)
    bad[:] = 0
    emul.emul_colblock_split(n_own, p(loffs), p(lcols), p(M.vals[rows]), p(cuts2), nb2, n_own, n_low, p(cnt2),
                             None, None, None, 0, p(bad))
    assert bad[0] == 0                             # renumbered rows are still in global order
    ordv = np.where(lcols < n_own, n_low + lcols.astype(np.int64),
                    np.where(lcols.astype(np.int64) - n_own < n_low, lcols.astype(np.int64) - n_own, lcols))
    cnt2 = cnt2.reshape(nb2, n_own + 1)
    for i in range(n_own):
        o = ordv[int(loffs[i]):int(loffs[i + 1])]
        assert np.all(np.diff(o) > 0)
        want = [int(np.sum((o >= cuts2[b]) & (o < cuts2[b + 1]))) for b in range(nb2)]
        assert list(cnt2[:, i]) == want


def test_grouped_kernel_with_row_major_bins_on_the_emulator(emul):
    """What one launch per column range does (k_spmv_sell_grp): the short rows as SELL slices
    in a length-sorted order (row ids through the permutation, rows past the end of the list,
    empty slices), four slices per warp trip, and the rows of the two row-major bins as units of
    the same work list -- a CTA per "long" row, a warp per "vector" row.  y = A x and y += A x
    against the CSR product to the 1e-13 bar; the SELL rows bit for bit (they add left to right)."""
    emul.emul_sell_grp_bins.argtypes = ([C.c_int, C.c_uint, C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32]
                                        + [C.c_uint32, C.c_void_p, C.c_void_p] * 2 + [C.c_void_p] * 2)
    n = 2500
    M = orc.gen_powerlaw(n, 4)
    lens = M.rowlens()
    offs = M.offs.astype(np.int64)
    sell_rows = np.flatnonzero(lens <= 12)
    vec_rows = np.flatnonzero((lens > 12) & (lens <= 300))
    long_rows = np.flatnonzero(lens > 300)
    assert len(vec_rows) > 20 and len(long_rows) > 2
    # SELL list: sorted by decreasing length inside windows of 256 rows, padded to 32, last slices empty
    order = np.concatenate([w[np.argsort(-lens[w], kind="stable")] for w in np.array_split(sell_rows, 8)])
    ns = (len(order) + 31) // 32 + 3                 # three empty slices at the end of the list
    perm = np.full(ns * 32 + 1, 0xFFFFFFFF, dtype=np.uint32)
    perm[:len(order)] = order
    sell_off, cols, vals = [0], [], []
    for s_ in range(ns):
        rows = perm[32 * s_:32 * s_ + 32]
        real = rows[rows != 0xFFFFFFFF]
        w = int(lens[real].max()) if len(real) else 0
        Cc, V = np.zeros((w, 32), dtype=np.uint32), np.zeros((w, 32))
        for l, r in enumerate(rows):
            if r == 0xFFFFFFFF:
                continue
            a, b = offs[r], offs[r + 1]
            Cc[:b - a, l], V[:b - a, l] = M.cols[a:b], M.vals[a:b]
            Cc[b - a:, l] = M.cols[b - 1]            # padding: the row's last column, value 0
        cols += Cc.reshape(-1).tolist()
        vals += V.reshape(-1).tolist()
        sell_off.append(sell_off[-1] + w)
    sell_off = np.array(sell_off, dtype=np.uint32)
    cols, vals = np.array(cols + [0] * 32, dtype=np.uint32), np.array(vals + [0.0] * 32)

    def row_major(rows, start):
        o, cc, vv = [start], [], []
        for r in rows:
            a, b = offs[r], offs[r + 1]
            pad = (-(b - a)) % 4
            cc += M.cols[a:b].tolist() + [int(r)] * pad
            vv += M.vals[a:b].tolist() + [0.0] * pad
            o.append(o[-1] + (b - a) + pad)
        return np.array(o, dtype=np.uint64), cc, vv
    voff, vc, vv = row_major(vec_rows, 0)
    loff, lc, lv = row_major(long_rows, int(voff[-1]))     # the long rows live after the vector rows
    vl_cols = np.array(vc + lc + [0] * 8, dtype=np.uint32)
    vl_vals = np.array(vv + lv + [0.0] * 8)
    vids, lids = vec_rows.astype(np.uint32), long_rows.astype(np.uint32)
    p = lambda a: a.ctypes.data
    x = np.random.default_rng(8).standard_normal(n)
    ref, scale = orc.spmv(M, x, want_abs=True)
    fma = orc.spmv_fma(M, x)
    for grid in (1, 3):
        y = np.full(n, np.nan)
        assert emul.emul_sell_grp_bins(0, grid, ns, p(sell_off), p(cols), p(vals), p(perm), p(x), p(y), n,
                                       len(lids), p(lids), p(loff), len(vids), p(vids), p(voff), p(vl_cols),
                                       p(vl_vals)) == 0
        assert np.all(np.abs(y - ref) <= 1e-13 * np.maximum(scale, 1e-300))
        assert y[sell_rows].tobytes() == fma[sell_rows].tobytes()
        y2 = y.copy()
        assert emul.emul_sell_grp_bins(1, grid, ns, p(sell_off), p(cols), p(vals), p(perm), p(x), p(y2), n,
                                       len(lids), p(lids), p(loff), len(vids), p(vids), p(voff), p(vl_cols),
                                       p(vl_vals)) == 0
        assert np.all(np.abs(y2 - 2.0 * ref) <= 2e-13 * np.maximum(scale, 1e-300))
