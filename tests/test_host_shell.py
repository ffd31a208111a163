"""The C host shell around the backend (lsbench_b200/host): public API, CLI,
reader, registration of `--solver b200` -- everything that runs without a GPU.

Anchors: src/lsbench.h:31-40 (API), src/lsbench.c:82-150 (CLI), :156-187
(dispatch), src/lsbench-csr.c:29-92 (reader), bin/driver.c (must build
unchanged), BASELINE.json config 1 (cholmod plumbing on CPU).
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
READER = json.load(open(os.path.join(GOLD, "reader.json")))


class Csr(C.Structure):
    _fields_ = [("nrows", C.c_uint), ("base", C.c_uint), ("offs", C.POINTER(C.c_uint)),
                ("cols", C.POINTER(C.c_uint)), ("vals", C.POINTER(C.c_double)),
                ("gen_kind", C.c_int), ("gen_size", C.c_ulonglong),
                ("gen_seed", C.c_ulonglong), ("gen_nnz", C.c_ulonglong)]


class Lsbench(C.Structure):
    _fields_ = [("matrix", C.c_char_p), ("solver", C.c_int), ("ordering", C.c_int),
                ("precision", C.c_int), ("verbose", C.c_uint), ("trials", C.c_uint)]


@pytest.fixture(scope="module")
def host():
    from lsbench_b200 import build, build_host
    build.build()
    build_host.build()
    L = C.CDLL(build_host.LIB)
    L.lsbench_matrix_read.restype = C.POINTER(Csr)
    L.lsbench_matrix_read.argtypes = [C.c_char_p]
    L.lsbench_matrix_free.argtypes = [C.POINTER(Csr)]
    L.lsbench_init.restype = C.POINTER(Lsbench)
    L.lsbench_init.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.lsbench_finalize.argtypes = [C.POINTER(Lsbench)]
    L.lsbench_get_matrix_name.restype = C.c_char_p
    L.lsbench_get_matrix_name.argtypes = [C.POINTER(Lsbench)]
    L.lsbench_matrix_nnz.restype = C.c_ulonglong
    L.lsbench_matrix_nnz.argtypes = [C.POINTER(Csr)]
    return L, build_host


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def init(L, *words):
    argv = (C.c_char_p * (len(words) + 2))(b"driver", *[w.encode() for w in words], None)
    return L.lsbench_init(len(words) + 1, argv)


def test_public_api_is_the_reference_api(host):
    L, _ = host
    for f in ("lsbench_matrix_read", "lsbench_matrix_print", "lsbench_matrix_free",
              "lsbench_init", "lsbench_get_matrix_name", "lsbench_bench", "lsbench_finalize",
              "b200_init", "b200_bench", "b200_finalize"):
        assert hasattr(L, f), f
    # the enumerators carry the reference's values (src/lsbench.h:8-29): ask the compiler
    want = {"LSBENCH_SOLVER_NONE": -1, "LSBENCH_SOLVER_CUSOLVER": 0, "LSBENCH_SOLVER_HYPRE": 1,
            "LSBENCH_SOLVER_AMGX": 2, "LSBENCH_SOLVER_CHOLMOD": 3, "LSBENCH_SOLVER_PARALMOND": 4,
            "LSBENCH_SOLVER_GINKGO": 5, "LSBENCH_SOLVER_B200": 6,
            "LSBENCH_PRECISION_FP64": 0, "LSBENCH_PRECISION_FP32": 1, "LSBENCH_PRECISION_FP16": 2,
            "LSBENCH_ORDERING_NONE": -1, "LSBENCH_ORDERING_RCM": 0, "LSBENCH_ORDERING_AMD": 1,
            "LSBENCH_ORDERING_METIS": 2}
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "enums.c")
        with open(src, "w") as f:
            f.write('#include "lsbench.h"\n#include <stdio.h>\nint main(void) {\n'
                    + "".join('  printf("%s %%d\\n", (int)%s);\n' % (k, k) for k in want)
                    + "  return 0;\n}\n")
        exe = os.path.join(d, "enums")
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = {l.split()[0]: int(l.split()[1]) for l in out.splitlines()}
    assert got == want


@pytest.mark.parametrize("name", orc.TOY + orc.NEK)
def test_host_reader_is_bit_identical_to_the_reference(host, name):
    L, _ = host
    p = L.lsbench_matrix_read(orc.matrix_path(name).encode())
    a, g = p.contents, READER[name]
    nnz = a.offs[a.nrows]
    assert (a.nrows, nnz, a.base) == (g["n"], g["nnz"], g["base"])
    assert sha(np.ctypeslib.as_array(a.offs, (a.nrows + 1,))) == g["offs"]
    assert sha(np.ctypeslib.as_array(a.cols, (nnz,))) == g["cols"]
    assert sha(np.ctypeslib.as_array(a.vals, (nnz,))) == g["vals"]
    assert L.lsbench_matrix_nnz(p) == nnz and a.gen_kind == 0
    L.lsbench_matrix_free(p)


def test_host_reader_sorts_folds_and_compresses(host, tmp_path):
    L, _ = host
    f = tmp_path / "m.txt"
    f.write_text("6 1\n3 1 5.0\n1 2 1.5\n1 1 2.0\n1 2 0.25\n4 4 7\n3 3 -1e0\n")
    p = L.lsbench_matrix_read(str(f).encode())
    a = p.contents
    assert a.nrows == 3 and [a.offs[i] for i in range(4)] == [0, 2, 4, 5]
    assert [a.cols[i] for i in range(5)] == [1, 2, 1, 3, 4]
    assert [a.vals[i] for i in range(5)] == [2.0, 1.75, 5.0, -1.0, 7.0]
    L.lsbench_matrix_free(p)


def test_synthetic_descriptor(host):
    L, _ = host
    p = L.lsbench_matrix_read(b"poisson27:512")
    a = p.contents
    assert (a.nrows, a.gen_kind, a.gen_size) == (512 ** 3, 2, 512) and not a.offs
    L.lsbench_matrix_free(p)
    p = L.lsbench_matrix_read(b"powerlaw:50000000:3")
    assert (p.contents.nrows, p.contents.gen_kind, p.contents.gen_seed) == (50000000, 3, 3)
    L.lsbench_matrix_free(p)


def test_cli_defaults_and_both_option_forms(host):
    L, _ = host
    cb = init(L, "--matrix", "m.txt")
    c = cb.contents  # defaults: src/lsbench.c:95-96
    assert (c.solver, c.ordering, c.precision, c.verbose, c.trials) == (0, 0, 0, 0, 100)
    assert L.lsbench_get_matrix_name(cb) == b"m.txt"
    L.lsbench_finalize(cb)
    cb = init(L, "--test", "t.txt", "--solver", "b200", "--trials", "7", "--verbose=2",
              "--ordering", "amd")
    c = cb.contents
    assert (c.matrix, c.solver, c.trials, c.verbose, c.ordering) == (b"t.txt", 6, 7, 2, 1)
    L.lsbench_finalize(cb)
    cb = init(L, "--matrix", "m", "--solver", "CuSolver", "--trials=3")
    assert (cb.contents.solver, cb.contents.trials) == (0, 3)
    L.lsbench_finalize(cb)
    cb = init(L, "--matrix", "m", "--solver", "cusparse")  # src/lsbench.c:31-33
    assert cb.contents.solver == 3
    L.lsbench_finalize(cb)


def test_fp32_is_accepted_for_b200_only(host):
    L, _ = host
    cb = init(L, "--matrix", "m", "--solver", "b200", "--precision=fp32")
    assert (cb.contents.solver, cb.contents.precision) == (6, 1)
    L.lsbench_finalize(cb)


def test_driver_cli_errors(host):
    _, bh = host
    r = subprocess.run([bh.DRIVER, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("Usage: " + bh.DRIVER)
    r = subprocess.run([bh.DRIVER, "--solver", "b200"], capture_output=True, text=True)
    assert r.returncode == 1 and "Input matrix file not provided" in r.stderr
    r = subprocess.run([bh.DRIVER, "--matrix", "x", "--precision=FP32"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "FP64" in r.stderr
    # FP32 has a meaning for b200 only (fp32-stored operator); FP16 for nobody
    for words in (["--solver", "cholmod", "--precision=FP32"], ["--solver", "b200", "--precision=FP16"]):
        r = subprocess.run([bh.DRIVER, "--matrix", "x"] + words, capture_output=True, text=True)
        assert r.returncode == 1 and "Precisions other than FP64 are not implemented yet." in r.stderr
    r = subprocess.run([bh.DRIVER, "--matrix", "/nonexistent.txt", "--solver", "b200", "--precision=FP32"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "Unable to open file" in r.stderr   # got past the precision check
    r = subprocess.run([bh.DRIVER, "--matrix", "/nonexistent.txt", "--solver", "b200"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "Unable to open file" in r.stderr
    r = subprocess.run([bh.DRIVER, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1


def test_reference_driver_source_builds_unchanged(host):
    _, bh = host
    if not os.path.exists(bh.DRIVER_REF):
        pytest.skip("reference tree not mounted when the host shell was built")
    r = subprocess.run([bh.DRIVER_REF, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--solver" in r.stdout


def test_cpu_only_host_still_runs_the_cholmod_configuration(host):
    """BASELINE.json config 1.  In the product library cholmod is 'not built'
    (like the reference with ENABLE_CHOLMOD=OFF) and the run must survive on a
    GPU-less host even though b200 is compiled in."""
    _, bh = host
    r = subprocess.run([bh.DRIVER, "--solver", "cholmod", "--test",
                        orc.matrix_path("I1_05x05")], capture_output=True, text=True)
    assert r.returncode == 0 and "did not run" in r.stderr


def test_config1_with_the_direct_solve_standin(host, tmp_path):
    """Same command with the test-only CHOLMOD stand-in linked in: CSV row in
    the reference's format and the analytic I1 solution."""
    _, bh = host
    from lsbench_b200 import build_host as b
    orc.lib()
    lib = str(tmp_path / "liblsbench.so")
    drv = str(tmp_path / "driver")
    host_dir = os.path.join(ROOT, "lsbench_b200", "host")
    inc = ["-I", os.path.join(ROOT, "include"), "-I", host_dir, "-I", orc.ORACLE_DIR]
    subprocess.run([b._cc(), "-O2", "-std=c11", "-fPIC", "-shared", "-DLSBENCH_CHOLMOD_STANDIN",
                    "-o", lib] + inc +
                   [os.path.join(host_dir, f) for f in ("lsbench.c", "lsbench-csr.c", "b200.c")] +
                   [os.path.join(ROOT, "tests", "cholmod_standin.c"), orc.LIB,
                    "-Wl,-rpath," + os.path.dirname(orc.LIB)], check=True)
    subprocess.run([b._cc(), "-O2", "-std=c11"] + inc + [os.path.join(host_dir, "main.c"), "-o", drv,
                    lib, "-Wl,-rpath," + str(tmp_path)], check=True)
    out = str(tmp_path / "x.bin")
    r = subprocess.run([drv, "--solver", "cholmod", "--test", orc.matrix_path("I1_05x05"),
                        "--trials=4", "--dump-x", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "===matrix,n,nnz,trials,solver,ordering,elapsed==="
    f = lines[1].split(",")
    assert f[0].endswith("I1_05x05.txt") and f[1:6] == ["5", "5", "4", "3", "0"]
    assert float(f[6]) >= 0
    np.testing.assert_allclose(np.fromfile(out), [0, 1 / 2, 2 / 3, 3 / 4, 4 / 5], rtol=1e-15)


def test_binary_cache_round_trip(host, tmp_path, monkeypatch):
    """LSBENCH_MATRIX_CACHE: the cached CSR is byte-identical to the parsed one,
    is written beside the text file, and is dropped when the text is newer
    (reader semantics src/lsbench-csr.c:29-92)."""
    L, _ = host
    f = tmp_path / "m.txt"
    f.write_text("6 1\n3 1 5.0\n1 2 1.5\n1 1 2.0\n1 2 0.25\n4 4 7\n3 3 -1e0\n")

    def grab():
        p = L.lsbench_matrix_read(str(f).encode())
        a = p.contents
        nnz = a.offs[a.nrows]
        out = (a.nrows, a.base, [a.offs[i] for i in range(a.nrows + 1)],
               [a.cols[i] for i in range(nnz)], [a.vals[i] for i in range(nnz)])
        L.lsbench_matrix_free(p)
        return out

    plain = grab()
    assert not os.path.exists(str(f) + ".lsbcsr")
    monkeypatch.setenv("LSBENCH_MATRIX_CACHE", "1")
    assert grab() == plain and os.path.exists(str(f) + ".lsbcsr")
    # second read comes from the cache: make the text unreadable as a matrix,
    # keep its mtime older than the cache
    st = os.stat(str(f))
    f.write_text("garbage\n")
    os.utime(str(f), (st.st_atime, st.st_mtime - 10))
    assert grab() == plain
    # a text file newer than the cache wins
    f.write_text("1 0\n0 0 3.5\n")
    os.utime(str(f), (st.st_atime + 100, st.st_mtime + 100))
    assert grab() == (1, 0, [0, 1], [0], [3.5])
    # a truncated cache is ignored, not trusted
    with open(str(f) + ".lsbcsr", "r+b") as g:
        g.truncate(20)
    os.utime(str(f) + ".lsbcsr", (st.st_atime + 200, st.st_mtime + 200))
    assert grab() == (1, 0, [0, 1], [0], [3.5])


@pytest.mark.parametrize("name", ["tj7a_A_18"])
def test_binary_cache_real_matrix(host, name, tmp_path, monkeypatch):
    L, _ = host
    import shutil
    f = str(tmp_path / (name + ".txt"))
    shutil.copy(orc.matrix_path(name), f)
    monkeypatch.setenv("LSBENCH_MATRIX_CACHE", "1")
    for _ in range(2):   # parse + store, then load
        p = L.lsbench_matrix_read(f.encode())
        a = p.contents
        nnz = a.offs[a.nrows]
        got = {"n": a.nrows, "base": a.base, "nnz": int(nnz),
               "offs": sha(np.ctypeslib.as_array(a.offs, (a.nrows + 1,))),
               "cols": sha(np.ctypeslib.as_array(a.cols, (nnz,))),
               "vals": sha(np.ctypeslib.as_array(a.vals, (nnz,)))}
        L.lsbench_matrix_free(p)
        for k, v in got.items():
            assert READER[name][k] == v, k


def _random_coo_text(rng):
    """a COO file the reference's fscanf grammar accepts (src/lsbench-csr.c:37,50):
    any white space before a field, the newline right after the value; records in
    any order, with duplicates and absent row ids"""
    base = int(rng.integers(0, 2))
    nrec = int(rng.integers(1, 60))
    rows = rng.integers(base, base + 12, nrec) * int(rng.integers(1, 4))
    cols = rng.integers(base, base + 15, nrec)
    fmts = ["%.17g", "%.6f", "%.3e", "%d", "%+.10e", "%.15f"]
    out = ["%d %d\n" % (nrec, base)]
    for r, c in zip(rows, cols):
        v = float(rng.standard_normal() * 10.0 ** int(rng.integers(-6, 7)))
        f = fmts[int(rng.integers(0, len(fmts)))]
        val = f % (int(v) if f == "%d" else v)
        ws = lambda: rng.choice([" ", "  ", "\t", " \t "])
        lead = rng.choice(["", "", "", " ", "\n", "\n\n  ", "\t"])
        out.append("%s%d%s%d%s%s\n" % (lead, r, ws(), c, ws(), val))
    return "".join(out)


def test_reader_fuzz_against_the_reference_and_the_oracle(host, tmp_path):
    """300 generated files: this tree's reader, the CPU oracle and -- when
    oracle/_ref is built -- the reference's own reader give the same CSR, bit for
    bit (duplicates are summed in file order by all three)."""
    L, _ = host
    rng = np.random.default_rng(2024)
    f = str(tmp_path / "fuzz.txt")
    have_ref = os.path.exists(orc.REF_LIB)
    for _ in range(300):
        text = _random_coo_text(rng)
        with open(f, "w") as g:
            g.write(text)
        want = orc.matrix_read(f)
        assert want is not None, text
        p = L.lsbench_matrix_read(f.encode())
        a = p.contents
        nnz = a.offs[a.nrows]
        got = (a.nrows, a.base, np.ctypeslib.as_array(a.offs, (a.nrows + 1,)).copy(),
               np.ctypeslib.as_array(a.cols, (nnz,)).copy(), np.ctypeslib.as_array(a.vals, (nnz,)).copy())
        L.lsbench_matrix_free(p)
        refs = [want] + ([orc.ref_matrix_read(f)] if have_ref else [])
        for w in refs:
            assert (got[0], got[1]) == (w.nrows, w.base), text
            assert np.array_equal(got[2], w.offs) and np.array_equal(got[3], w.cols), text
            assert got[4].tobytes() == w.vals.tobytes(), text


# ---- --ordering: host-side RCM, no GPU needed ---------------------

def _order_fns(L):
    L.b200_host_rcm.restype = C.c_int
    L.b200_host_rcm.argtypes = [C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    L.b200_host_reordered_operator.restype = C.POINTER(Csr)
    L.b200_host_reordered_operator.argtypes = [C.POINTER(Csr), C.c_int, C.POINTER(C.c_uint)]


def _bandwidth(S):
    S = S.tocoo()
    return int(np.abs(S.row.astype(np.int64) - S.col).max()) if S.nnz else 0


def _rcm(L, S):
    S = S.tocsr()
    S.sort_indices()
    offs = np.ascontiguousarray(S.indptr, dtype=np.uint32)
    cols = np.ascontiguousarray(S.indices, dtype=np.uint32)
    perm = np.zeros(max(S.shape[0], 1), dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint))
    assert L.b200_host_rcm(S.shape[0], p(offs), p(cols), p(perm)) == 0
    return perm[:S.shape[0]]


@pytest.mark.parametrize("name", ["tj7a_A_12", "tj7a_A_15", "tj7a_A_18", "xn3b_A_10", "xn3b_A_12",
                                  "xn3b_A_15", "xn3b_A_18"])
def test_rcm_is_a_permutation_and_as_narrow_as_scipy(host, name):
    """RCM of the operator that is solved: a permutation, deterministic, and a
    bandwidth comparable with scipy's reverse_cuthill_mckee (measured here, file
    numbering -> this code / scipy: xn3b_A_10 2343 -> 1053 / 1053, xn3b_A_12
    1975 -> 961 / 724, xn3b_A_18 2123 -> 568 / 772; the tj7a files are already
    numbered better than either, 277 -> 443 / 444)"""
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    L, _ = host
    _order_fns(L)
    S = orc.op_upper_mirror(orc.matrix_read(orc.matrix_path(name))).scipy()
    perm = _rcm(L, S)
    assert np.array_equal(np.sort(perm), np.arange(S.shape[0]))
    assert np.array_equal(perm, _rcm(L, S))
    mine = _bandwidth(S[perm][:, perm])
    sp = reverse_cuthill_mckee(S.tocsr(), symmetric_mode=True)
    theirs = _bandwidth(S[sp][:, sp])
    assert mine <= 1.4 * theirs, (mine, theirs, _bandwidth(S))
    if name.startswith("xn3b"):
        assert mine < 0.55 * _bandwidth(S)


def test_rcm_edge_shapes(host):
    """one row; a diagonal matrix (n components); two components; an
    unsymmetric pattern (ordered on S + S^T); a path numbered at random comes
    back with bandwidth 1"""
    import scipy.sparse as sp
    L, _ = host
    _order_fns(L)
    assert _rcm(L, sp.identity(1, format="csr")).tolist() == [0]
    assert np.array_equal(np.sort(_rcm(L, sp.identity(7, format="csr"))), np.arange(7))
    two = sp.block_diag([sp.diags([1.0, 1.0], [0, 1], shape=(4, 4)), sp.diags([1.0, 1.0], [0, 1], shape=(3, 3))]).tocsr()
    p = _rcm(L, two)
    assert np.array_equal(np.sort(p), np.arange(7)) and _bandwidth((two + two.T).tocsr()[p][:, p]) == 1
    rng = np.random.default_rng(5)
    n = 200
    q = rng.permutation(n)
    path = sp.diags([1.0, 2.0], [-1, 0], shape=(n, n)).tocsr()[q][:, q]   # lower bidiagonal, shuffled
    p = _rcm(L, path)
    assert _bandwidth(path[p][:, p]) == 1


@pytest.mark.parametrize("name,sym", [("xn3b_A_15", 1), ("xn3b_A_15", 0), ("I1_05x05", 1), ("A1_02x02", 0)])
def test_reordered_operator_is_P_S_Pt(host, name, sym):
    """what b200_bench hands to the conversion under an ordering: the operator
    that is solved (upper triangle mirrored like src/cholmod-impl.h:5-21, or the
    matrix as stored) with rows and columns renumbered, values bit for bit,
    columns ascending"""
    L, _ = host
    _order_fns(L)
    A = orc.matrix_read(orc.matrix_path(name))
    S = (orc.op_upper_mirror(A) if sym else orc.op_full(A)).scipy().tocsr()
    a = L.lsbench_matrix_read(orc.matrix_path(name).encode())
    perm = np.zeros(A.nrows, dtype=np.uint32)
    b = L.b200_host_reordered_operator(a, sym, perm.ctypes.data_as(C.POINTER(C.c_uint)))
    B = b.contents
    n, nnz = B.nrows, B.offs[B.nrows]
    got = (int(B.base), np.ctypeslib.as_array(B.offs, (n + 1,)).copy(),
           np.ctypeslib.as_array(B.cols, (nnz,)).copy(), np.ctypeslib.as_array(B.vals, (nnz,)).copy())
    L.lsbench_matrix_free(b)
    L.lsbench_matrix_free(a)
    assert np.array_equal(np.sort(perm), np.arange(n))
    want = S[perm][:, perm].tocsr()
    want.sort_indices()
    assert got[0] == 0 and np.array_equal(got[1], want.indptr) and np.array_equal(got[2], want.indices)
    assert got[3].tobytes() == want.data.tobytes()


# ---- the drop-in, done for real: --solver b200 inside the stock lsbench tree ---------------
DROPIN = os.path.join(ROOT, "oracle", "_ref", "driver_ref_b200")


def test_b200_drops_into_the_stock_lsbench_tree(host):
    """oracle/dropin.py copies the reference tree to a scratch directory, applies
    the five registration edits of INTEGRATION.md, adds this repository's b200.c
    (unchanged) and builds the reference's own sources with -DLSBENCH_B200: the
    stock `driver` then knows `--solver b200`, fails loudly when there is no GPU
    (no fallback), and its other backends behave as before"""
    if not os.path.isdir("/root/reference/src"):
        if not os.path.exists(DROPIN):
            pytest.skip("no reference tree and no prebuilt oracle/_ref/driver_ref_b200")
    else:
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "dropin.py")], check=True,
                       capture_output=True)
    m = orc.matrix_path("I1_05x05")
    r = subprocess.run([DROPIN, "--solver", "cholmod", "--matrix", m, "--trials=2"], capture_output=True, text=True)
    assert r.returncode == 0                        # the stock stub: returns without a solve
    r = subprocess.run([DROPIN, "--solver", "b200", "--matrix", m, "--trials=2"], capture_output=True, text=True)
    from lsbench_b200 import abi
    try:
        have_gpu = abi.device_count() > 0
    except abi.B200Error:
        have_gpu = False
    if have_gpu:
        assert r.returncode == 0 and "===b200:" in r.stdout
    else:
        assert r.returncode == 1 and "b200 error" in r.stderr and "cuda error" in r.stderr
