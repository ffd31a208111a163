"""The SELL SpMV kernels of the product, compiled for the HOST by
tests/spmv_emul.cpp and run thread by thread over a launch grid: indexing and
the order of the additions against the oracle's fma CSR product, bit for bit,
without a GPU.  Covered: the default kernels k_spmv_sellc / k_spmv_sell
(lsbench_b200/csrc/sell_kernels.cuh; fp64 and fp32 value streams).  The index-compressed SELL layout (DESIGN.md
section 2, csrc/convert.cu k_sell_fill / k_slice_uniform / k_compact_cols) is
restated here with numpy."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NONE = 0xFFFFFFFF


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libspmv_emul.so")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC",
                    "-I", os.path.join(ROOT, "lsbench_b200", "csrc"),
                    os.path.join(ROOT, "tests", "spmv_emul.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.emul_sellc.argtypes = [C.c_int, C.c_uint] + [C.c_void_p] * 7 + [C.c_uint32] * 5
    L.emul_sellc9.argtypes = [C.c_uint] + [C.c_void_p] * 7 + [C.c_uint32] * 5
    L.emul_sell.argtypes = [C.c_int, C.c_uint] + [C.c_void_p] * 6 + [C.c_uint32] * 5
    return L


def sellc_layout(M, perm=None):
    """index-compressed SELL-32 with fp32 values of the operator M (orc.Op)"""
    n = M.n
    lens = M.rowlens()
    ns = (n + 31) // 32
    lst = np.full(ns * 32 + 1, NONE, dtype=np.uint32)
    lst[:n] = np.arange(n) if perm is None else perm
    meta = np.zeros((ns + 1, 4), dtype=np.uint32)
    vals, ecols, dcols, allcols, sell_off = [], [], [], [], [0]
    o = 0
    for s in range(ns):
        rows = lst[32 * s:32 * s + 32]
        real = rows != NONE
        w = int(lens[rows[real]].max()) if real.any() else 0
        V = np.zeros((w, 32), dtype=np.float32)
        Cc = np.zeros((w, 32), dtype=np.uint32)
        for l, r in enumerate(rows):
            if r == NONE:
                continue
            a, b = int(M.offs[r]), int(M.offs[r + 1])
            V[:b - a, l] = M.vals[a:b]
            Cc[:b - a, l] = M.cols[a:b]
            Cc[b - a:, l] = r                        # padding: own row, value 0
        uniform = bool(real.all() and w > 0 and (lens[rows] == w).all()
                       and all(len(set((Cc[k].astype(np.int64) - rows.astype(np.int64)).tolist())) == 1
                               for k in range(w)))
        if uniform:
            meta[s] = (o, w | 0x80000000, len(dcols), 0)
            dcols += (Cc[:, 0].astype(np.int64) - int(rows[0])).tolist()
        else:
            meta[s] = (o, w, len(ecols) // 32, 0)
            ecols += Cc.reshape(-1).tolist()
        vals += V.reshape(-1).tolist()
        allcols += Cc.reshape(-1).tolist()
        o += w
        sell_off.append(o)
    return dict(ns=ns, meta=meta, list=None if perm is None else lst,
                vals=np.array(vals + [0.0], dtype=np.float32),
                ecols=np.array(ecols + [0], dtype=np.uint32),
                dcols=np.array(dcols + [0] * 40, dtype=np.int32),
                allcols=np.array(allcols + [0], dtype=np.uint32), sell_off=np.array(sell_off, dtype=np.uint32),
                uniform=int((meta[:ns, 1] >> 31).sum()), wmax=int((meta[:ns, 1] & 0x7FFFFFFF).max()))


def run(emul, Lay, n, x, wmax, grid, ranges=None):
    """every kernel on the same layout -- k_spmv_sellc and k_spmv_sell, each with both
    value types; the answers must not differ"""
    b0, e0, b1, e1 = ranges or (0, Lay["ns"], 0, 0)
    p = lambda a: None if a is None else a.ctypes.data
    ys = []
    y = np.full(n, np.nan)    # chunks of 9 (what the product runs on 27-wide rows), any width
    assert emul.emul_sellc9(grid, p(Lay["meta"]), p(Lay["ecols"]), p(Lay["dcols"]),
                            p(Lay["vals"].astype(np.float64)), p(Lay["list"]), p(x), p(y), b0, e0, b1, e1, n) == 0
    ys.append(y)
    # the default kernels on the same layout: index-compressed and explicit columns
    for f64, vals in ((0, Lay["vals"]), (1, Lay["vals"].astype(np.float64))):
        y = np.full(n, np.nan)
        assert emul.emul_sellc(f64, grid, p(Lay["meta"]), p(Lay["ecols"]), p(Lay["dcols"]), p(vals),
                               p(Lay["list"]), p(x), p(y), b0, e0, b1, e1, n) == 0
        ys.append(y)
        y = np.full(n, np.nan)
        assert emul.emul_sell(f64, grid, p(Lay["sell_off"]), p(Lay["allcols"]), p(vals),
                              p(Lay["list"]), p(x), p(y), b0, e0, b1, e1, n) == 0
        ys.append(y)
    assert all(y.tobytes() == ys[0].tobytes() for y in ys)
    return ys[0]


@pytest.mark.parametrize("gen,N,wmax", [("poisson27", 8, 32), ("poisson7", 12, 8), ("poisson7", 12, 16),
                                        ("poisson27", 20, 32)])
def test_sell_kernels_on_stencils(emul, gen, N, wmax):
    M = getattr(orc, "gen_" + gen)(N)
    assert np.array_equal(M.vals.astype(np.float32).astype(np.float64), M.vals)
    Lay = sellc_layout(M)
    x = np.random.default_rng(1).standard_normal(M.n)
    want = orc.spmv_fma(M, x)
    for grid in (1, 3, 148 * 2):
        assert run(emul, Lay, M.n, x, wmax, grid).tobytes() == want.tobytes()
    # the two-range form the overlapped multi-GPU SpMV uses (interior, then boundary slices)
    ns = Lay["ns"]
    y = run(emul, Lay, M.n, x, wmax, 5, (ns // 4, ns // 2, 0, 0))
    lo, hi = 32 * (ns // 4), min(32 * (ns // 2), M.n)
    assert y[lo:hi].tobytes() == want[lo:hi].tobytes() and np.isnan(y[:lo]).all() and np.isnan(y[hi:]).all()
    y2 = run(emul, Lay, M.n, x, wmax, 5, (0, ns // 4, ns // 2, ns))
    assert np.isnan(y2[lo:hi]).all() and y2[:lo].tobytes() == want[:lo].tobytes()
    assert y2[hi:].tobytes() == want[hi:].tobytes()


@pytest.mark.parametrize("gen,wmax", [("poisson27", 32), ("poisson7", 8)])
def test_sell_kernels_uniform_slices_of_a_row_block(emul, gen, wmax):
    """40 x-lines out of the middle of a 96^3 grid, as a rank of the row-block
    partition holds them (global column ids): the slice in the middle of every
    x-line is uniform (w deltas instead of 32 w columns), the two with a line end
    are not -- both paths of the kernel, next to each other"""
    N = 96
    r0 = N * N * 5 + N * 7
    M = getattr(orc, "gen_" + gen)(N, r0, r0 + N * 40)
    Lay = sellc_layout(M)
    assert 0 < Lay["uniform"] < Lay["ns"] and Lay["wmax"] in (7, 27)
    x = np.random.default_rng(4).standard_normal(N ** 3)
    want = orc.spmv_fma(M, x)
    for grid in (1, 7):
        assert run(emul, Lay, M.n, x, wmax, grid).tobytes() == want.tobytes()


def test_sell_kernels_ragged_rows_and_a_permuted_list(emul):
    """rows of every length 1..32 (explicit slices, every tail length), the last
    slice half empty, with the identity list and with a length-sorted list"""
    rng = np.random.default_rng(7)
    n = 32 * 37 + 13
    lens = (np.arange(n) % 32) + 1
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    cols = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]).astype(np.uint32)
    vals = rng.integers(-1000, 1000, int(offs[-1])).astype(np.float64) / 64.0
    M = orc.Op(n, offs, cols, vals)
    x = rng.standard_normal(n)
    want = orc.spmv_fma(M, x)
    for perm in (None, np.argsort(-lens, kind="stable").astype(np.uint32)):
        Lay = sellc_layout(M, perm)
        assert Lay["wmax"] == 32
        for grid in (1, 4):
            assert run(emul, Lay, n, x, 32, grid).tobytes() == want.tobytes()


def test_sell_kernels_banded_rows_compress_in_any_numbering(emul):
    """a banded operator of constant row length: every full slice is uniform,
    and the deltas may be negative"""
    n, half = 32 * 9, 5
    rows = np.arange(n)
    offs = [0]
    cols, vals = [], []
    for i in rows:
        c = [(i + d) % n for d in range(-half, half + 1)]
        cols += c
        vals += [float(d * d + 1) if d else 64.0 for d in range(-half, half + 1)]
        offs.append(len(cols))
    M = orc.Op(n, np.array(offs, dtype=np.uint64), np.array(cols, dtype=np.uint32), np.array(vals))
    Lay = sellc_layout(M)
    assert 0 < Lay["uniform"] < Lay["ns"]            # the wrap-around slices are not uniform
    x = np.random.default_rng(3).standard_normal(n)
    assert run(emul, Lay, n, x, 16, 2).tobytes() == orc.spmv_fma(M, x).tobytes()


def test_default_kernels_on_wide_ragged_rows(emul):
    """rows of every length 1..70: full chunks of 8 (fp64 values) and 16 (fp32
    values) plus every tail length, in explicit slices, for k_spmv_sellc and
    k_spmv_sell """
    rng = np.random.default_rng(17)
    n = 70 * 12 + 5
    lens = (np.arange(n) % 70) + 1
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    cols = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]).astype(np.uint32)
    vals = rng.integers(-1000, 1000, int(offs[-1])).astype(np.float64) / 64.0
    M = orc.Op(n, offs, cols, vals)
    x = rng.standard_normal(n)
    want = orc.spmv_fma(M, x)
    for perm in (None, np.argsort(-lens, kind="stable").astype(np.uint32)):
        Lay = sellc_layout(M, perm)
        assert Lay["wmax"] == 70
        for grid in (1, 6):
            assert run(emul, Lay, n, x, 1, grid).tobytes() == want.tobytes()
