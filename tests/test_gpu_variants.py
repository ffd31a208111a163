"""GPU parity of the SURVEY 8(f) rows 2 and 4 variants, through the C ABI:

  B200_MAT_VALUES_F32        SELL values stored as fp32 -- lossless on the
                             stencils (same bits as the fp64-stored matrix),
                             fp64 refinement otherwise (same fp64 bars)
  B200_PCG_CHEBYSHEV2 / 3    Chebyshev-Jacobi on the on-chip coarse-grid kernel

Oracles: oracle/krylov.c orc_pcg_cheb / orc_pcg_refine32, the SuperLU direct
solves (tests/golden/direct.npz), and the default fp64 path of the same library.
"""
import os

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIRECT = np.load(os.path.join(GOLD, "direct.npz"))


@pytest.fixture(scope="module")
def abi():
    from lsbench_b200 import abi as m
    m.load()
    return m


@pytest.fixture(scope="module")
def ctx(abi):
    c = abi.Context(0)
    yield c
    c.close()


def op_to_csr(M):
    return orc.HostCsr(M.n, 0, M.offs.astype(np.uint32), M.cols, M.vals)


def make(abi, ctx, A, flags=0):
    return abi.Matrix.from_csr(ctx, A.nrows, A.base, A.offs, A.cols, A.vals, flags)


# ------------------------------------------------------------------ fp32-stored values
@pytest.mark.parametrize("gen,N", [("poisson7", 40), ("poisson27", 33), ("poisson27", 64)])
@pytest.mark.parametrize("compress", [True, False])
def test_f32_values_are_lossless_on_stencils(abi, ctx, gen, N, compress):
    """every stencil value is an fp32 number: the fp64 copy is dropped
    (values_f32 == 1), the export returns the same CSR bit for bit, SpMV has the
    bits of the oracle's fma product, and the PCG takes the same number of
    iterations to the same solution as the fp64-stored matrix"""
    M = getattr(orc, "gen_" + gen)(N)
    A = op_to_csr(M)
    base = 0 if compress else abi.MAT_NO_COMPRESS
    M64 = make(abi, ctx, A, base)
    M32 = make(abi, ctx, A, base | abi.MAT_VALUES_F32)
    i64, i32 = M64.info(), M32.info()
    assert (i64.values_f32, i32.values_f32) == (0, 1)
    assert i32.matrix_stream_bytes == i64.matrix_stream_bytes - 4 * i64.nnz_padded
    assert i32.device_bytes < i64.device_bytes
    offs, cols, vals = M32.export()
    assert np.array_equal(offs, M.offs) and np.array_equal(cols, M.cols)
    assert vals.tobytes() == M.vals.tobytes()
    x = np.random.default_rng(3).standard_normal(M.n)
    y32, y64 = M32.spmv_host(x), M64.spmv_host(x)
    assert y32.tobytes() == y64.tobytes() == orc.spmv_fma(M, x).tobytes()
    b = orc.rhs(M.n)
    xa, ra, _ = M64.pcg_host(b, flags=abi.PCG_NO_SMALL)
    xb, rb, _ = M32.pcg_host(b, flags=abi.PCG_NO_SMALL)
    assert (rb.status, rb.outer_iters) == (0, 0) and abs(rb.iters - ra.iters) <= 1
    assert np.linalg.norm(xb - xa) / np.linalg.norm(xa) <= 1e-10
    assert orc.true_relres(M, b, xb) <= 1e-10
    # run to run: identical bits
    xc, rc_, _ = M32.pcg_host(b, flags=abi.PCG_NO_SMALL)
    assert rc_.iters == rb.iters and xc.tobytes() == xb.tobytes()
    M64.close(), M32.close()


def test_f32_values_odd_widths(abi, ctx):
    """rows of every length 1..40 (chunks of 16 + a tail of 0..15) with exactly
    representable values: bit-identical to the oracle's fma product"""
    rng = np.random.default_rng(11)
    n = 4000
    lens = (np.arange(n) % 40) + 1
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    cols = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]).astype(np.uint32)
    vals = rng.integers(-1000, 1000, int(offs[-1])).astype(np.float64) / 64.0
    M = orc.Op(n, offs, cols, vals)
    for flags in (abi.MAT_VALUES_F32, abi.MAT_VALUES_F32 | abi.MAT_NO_SORT):
        Md = make(abi, ctx, op_to_csr(M), flags)
        assert Md.info().values_f32 == 1
        x = rng.standard_normal(n)
        assert Md.spmv_host(x).tobytes() == orc.spmv_fma(M, x).tobytes()
        Md.close()


@pytest.mark.parametrize("name", ["tj7a_A_18", "xn3b_A_10"])
def test_f32_values_rounded_then_refined(abi, ctx, name):
    """the Nek values do not survive the rounding: both streams stay
    (values_f32 == 2), b200_spmv still multiplies with the fp64 operator, and the
    solve is iterative refinement to the same bars -- true residual <= 1e-10 and x
    within 1e-8 of the direct solution -- with the oracle's pass count"""
    A = orc.matrix_read(orc.matrix_path(name))
    M = orc.op_upper_mirror(A)
    Md = make(abi, ctx, A, abi.MAT_SYM_UPPER | abi.MAT_VALUES_F32)
    assert Md.info().values_f32 == 2
    offs, cols, vals = Md.export()
    assert vals.tobytes() == M.vals.tobytes()
    x = np.random.default_rng(2).standard_normal(M.n)
    assert Md.spmv_host(x).tobytes() == orc.spmv_fma(M, x).tobytes()
    b = orc.rhs(M.n)
    xs, res, rc = Md.pcg_host(b, flags=abi.PCG_NO_SMALL)
    xo, ito, outo, relo, rco = orc.pcg_refine32(M, b)
    assert rc == 0 and res.status == 0 and res.true_relres <= 1e-10
    assert abs(res.outer_iters - outo) <= 1 and abs(res.iters - ito) <= 0.15 * ito, (res.iters, ito, res.outer_iters, outo)
    assert orc.true_relres(M, b, xs) <= 1e-10
    xg = DIRECT[name]
    assert np.linalg.norm(xs - xg) / np.linalg.norm(xg) <= 1e-8
    x2, r2, _ = Md.pcg_host(b, flags=abi.PCG_NO_SMALL)
    assert r2.iters == res.iters and x2.tobytes() == xs.tobytes()
    Md.close()


# ------------------------------------------------------------------ Chebyshev-Jacobi on the on-chip path
@pytest.mark.parametrize("name", ["tj7a_A_12", "tj7a_A_18", "xn3b_A_10", "xn3b_A_18"])
def test_chebyshev_jacobi_on_the_onchip_kernel(abi, ctx, name):
    """SURVEY 8(f) row 2, the preconditioner half (B200_PCG_CHEBYSHEV2 / 3): the on-chip
    coarse-grid kernel with a polynomial in D^-1 A as preconditioner against the
    oracle's statement of the method (orc_pcg_cheb: same interval rule, iteration counts
    within 2) and the direct solve (1e-8), at the 1e-10 bar on the true residual;
    reproducible bit for bit; degree 1 is the kernel as it was."""
    A = orc.matrix_read(orc.matrix_path(name))
    Mo = orc.op_upper_mirror(A)
    b = orc.rhs(Mo.n)
    M = make(abi, ctx, A, abi.MAT_SYM_UPPER)
    x1, r1, rc1 = M.pcg_host(b, tol=1e-10, maxit=5000)
    assert rc1 == 0 and r1.path == 1 and r1.outer_iters == 1
    for deg, fl in ((2, abi.PCG_CHEBYSHEV2), (3, abi.PCG_CHEBYSHEV3)):
        x, r, rc = M.pcg_host(b, tol=1e-10, maxit=5000, flags=fl)
        assert rc == 0 and r.status == 0 and r.path == 1 and r.outer_iters == deg
        assert r.true_relres <= 1e-10 and orc.true_relres(Mo, b, x) <= 1e-10
        assert np.linalg.norm(x - DIRECT[name]) / np.linalg.norm(DIRECT[name]) <= 1e-8
        _, ito, _, rco = orc.pcg_cheb(Mo, b, degree=deg)
        assert rco == 0 and abs(r.iters - ito) <= 2, (deg, r.iters, ito)
        assert r.iters < 0.62 * r1.iters
        x2, r2, _ = M.pcg_host(b, tol=1e-10, maxit=5000, flags=fl)
        assert r2.iters == r.iters and x2.tobytes() == x.tobytes()
    # the streaming path ignores the flag
    xs, rs, _ = M.pcg_host(b, tol=1e-10, maxit=5000, flags=abi.PCG_NO_SMALL | abi.PCG_CHEBYSHEV2)
    assert rs.path == 0 and rs.status == 0 and abs(rs.iters - r1.iters) <= 2
    M.close()


# ------------------------------------------------------------------ column blocking (config 5)
@pytest.mark.parametrize("kernel,sigma", [("grouped", "32768"), ("grouped", "1024"), ("plain", "0")])
def test_column_blocked_spmv(abi, ctx, kernel, sigma, monkeypatch):
    """B200_MAT_COL_BLOCK on the power-law operator (BASELINE.json config 5, 1.7 M rows, 13
    ranges of 1 MB of x): the layout exports the generator's CSR bit for bit, the SpMV meets
    the 1e-13 bar against the oracle's rows (row sums are formed range by range), agrees with
    the unblocked layout, and is bit-reproducible -- with the default kernel (four slices per
    warp trip, work handed out first come first served) and with the plain one."""
    monkeypatch.setenv("B200_COL_BLOCK_MB", "1")
    monkeypatch.setenv("B200_COL_BLOCK_KERNEL", kernel)
    monkeypatch.setenv("B200_COL_BLOCK_SIGMA", sigma)
    n = 1_700_000
    Mo = orc.gen_powerlaw(n, 1, 0, 4096)          # the oracle's first 4096 rows, global columns
    M = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, n, seed=1, flags=abi.MAT_COL_BLOCK)
    assert M.info().col_blocks == 13
    x = np.random.default_rng(0).standard_normal(n)
    y = M.spmv_host(x)
    ref, scale = orc.spmv(Mo, x, want_abs=True)
    assert np.all(np.abs(y[:4096] - ref) <= 1e-13 * np.maximum(scale, 1e-300))
    assert M.spmv_host(x).tobytes() == y.tobytes()
    M0 = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, n, seed=1, flags=0)
    y0 = M0.spmv_host(x)
    assert np.max(np.abs(y - y0)) <= 1e-11 * np.max(np.abs(y0))
    o0, c0, v0 = M0.export()
    o1, c1, v1 = M.export()
    assert np.array_equal(o0, o1) and np.array_equal(c0, c1) and v0.tobytes() == v1.tobytes()
    M.close(), M0.close()


def test_column_blocked_spmv_small_shapes(abi, ctx, monkeypatch):
    """ranges narrower than a sort window, a last range that is nearly empty, rows with no
    entry in a range (their y must survive the accumulate passes untouched), fewer slices
    than one unit of work"""
    monkeypatch.setenv("B200_COL_BLOCK_MB", "1")
    for n in (131_072 + 77, 300_001):
        Mo = orc.gen_powerlaw(n, 3)
        M = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, n, seed=3, flags=abi.MAT_COL_BLOCK)
        assert M.info().col_blocks == -(-n // 131072)
        x = np.random.default_rng(n).standard_normal(n)
        y = M.spmv_host(x)
        ref, scale = orc.spmv(Mo, x, want_abs=True)
        assert np.all(np.abs(y - ref) <= 1e-13 * np.maximum(scale, 1e-300))
        M.close()


# ------------------------------------------------------------------ block-Jacobi on the on-chip path
@pytest.mark.parametrize("name", orc.NEK)
def test_block_jacobi_on_the_onchip_kernel(abi, ctx, name):
    """SURVEY 8(f) row 2 (B200_PCG_BLOCK_JACOBI): the on-chip coarse-grid kernel with the
    inverted diagonal blocks of its row chunks as preconditioner, against the oracle's
    statement of the method on the partition the library reports (orc_pcg_bj: iteration
    counts within 3) and the direct solve (1e-8), at the 1e-10 bar on the true residual;
    fewer iterations than Jacobi; reproducible bit for bit; the streaming path ignores it."""
    A = orc.matrix_read(orc.matrix_path(name))
    Mo = orc.op_upper_mirror(A)
    b = orc.rhs(Mo.n)
    M = make(abi, ctx, A, abi.MAT_SYM_UPPER)
    part, bs = M.block_jacobi_partition()
    assert bs in (16, 32)
    sizes = np.bincount(part)
    assert sizes.max() == bs and np.sum(sizes < bs) <= 16      # one short block per row chunk at most
    x1, r1, _ = M.pcg_host(b, tol=1e-10, maxit=5000)
    assert r1.path == 1 and r1.block_jacobi == 0
    x, r, rc = M.pcg_host(b, tol=1e-10, maxit=5000, flags=abi.PCG_BLOCK_JACOBI)
    assert rc == 0 and r.status == 0 and r.path == 1 and r.block_jacobi == bs
    assert r.true_relres <= 1e-10 and orc.true_relres(Mo, b, x) <= 1e-10
    assert np.linalg.norm(x - DIRECT[name]) / np.linalg.norm(DIRECT[name]) <= 1e-8
    _, ito, _, rco = orc.pcg_bj(Mo, b, part)
    assert rco == 0 and abs(r.iters - ito) <= 3 + 8 * r.replacements, (r.iters, ito, r.replacements)
    assert r.iters < 0.9 * r1.iters
    x2, r2, _ = M.pcg_host(b, tol=1e-10, maxit=5000, flags=abi.PCG_BLOCK_JACOBI)
    assert r2.iters == r.iters and x2.tobytes() == x.tobytes()
    xs, rs, _ = M.pcg_host(b, tol=1e-10, maxit=5000, flags=abi.PCG_NO_SMALL | abi.PCG_BLOCK_JACOBI)
    assert rs.path == 0 and rs.status == 0 and rs.block_jacobi == 0 and abs(rs.iters - r1.iters) <= 2
    # x0 given, and the operator as stored (unsymmetric at 1e-8: the blocks are symmetrised)
    x3, r3, _ = M.pcg_host(b, x0=0.5 * x, tol=1e-10, maxit=5000, flags=abi.PCG_BLOCK_JACOBI)
    assert r3.status == 0 and r3.true_relres <= 1e-10
    M.close()
    Mf = make(abi, ctx, A, 0)
    xf, rf, rcf = Mf.pcg_host(b, tol=1e-10, maxit=5000, flags=abi.PCG_BLOCK_JACOBI)
    assert rcf == 0 and rf.status == 0 and rf.true_relres <= 1e-10 and rf.block_jacobi in (16, 32)
    Mf.close()


@pytest.mark.parametrize("flags_name", ["jacobi", "block_jacobi"])
def test_onchip_kernel_residual_replacement(abi, ctx, flags_name):
    """The exit check of the on-chip kernel (csrc/small.cu): at a bar close to what fp64 reaches
    the recurrence residual gets there before b - A x does; the kernel then replaces r by
    b - A x and goes on (status 0, the bar met on the TRUE residual) -- or, below what fp64
    can do, ends with status 4 instead of running to maxit.  Jacobi and block-Jacobi."""
    fl = abi.PCG_BLOCK_JACOBI if flags_name == "block_jacobi" else 0
    A = orc.matrix_read(orc.matrix_path("tj7a_A_18"))
    Mo = orc.op_upper_mirror(A)
    b = orc.rhs(Mo.n)
    M = make(abi, ctx, A, abi.MAT_SYM_UPPER)
    seen_replacement = False
    for tol in (1e-11, 4e-12, 2e-12, 1e-12):
        x, r, rc = M.pcg_host(b, tol=tol, maxit=5000, flags=fl)
        assert r.path == 1 and r.status in (0, 4), (tol, r.status)
        seen_replacement |= r.replacements > 0
        if r.status == 0:
            assert r.true_relres <= tol and orc.true_relres(Mo, b, x) <= 1.02 * tol, (tol, r.true_relres)
        assert np.linalg.norm(x - DIRECT["tj7a_A_18"]) / np.linalg.norm(DIRECT["tj7a_A_18"]) <= 1e-8
        assert r.iters < 1500                      # never to maxit
    x, r, rc = M.pcg_host(b, tol=3e-14, maxit=5000, flags=fl)
    assert r.status == 4 and r.iters < 1500 and r.replacements >= 1
    assert seen_replacement or r.replacements >= 1
    M.close()
