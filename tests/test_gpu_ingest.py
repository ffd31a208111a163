"""GPU parity for SURVEY 8(f) row 1, the device-side ingest: b200_coo_to_csr /
b200_mat_from_coo against the CPU oracle's reader (oracle/csr_read.c) and, when
oracle/_ref is present, the reference's own lsbench_matrix_read
(src/lsbench-csr.c:29-92), bit for bit.  Run on a B200: pytest -m gpu.

The inputs are the reference's matrices made hostile: records shuffled,
entries split into several records of the same (row, col) whose sum depends on
the order of addition, and row ids spread out so that most ids are absent.
"""
import os
import subprocess

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


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


def read_records(path):
    """(base, rows, cols, vals) of a COO text file, in file order."""
    with open(path) as f:
        nnz, base = (int(t) for t in f.readline().split())
        body = np.array(f.read().split())
    assert body.size == 3 * nnz
    body = body.reshape(nnz, 3)
    return (base, body[:, 0].astype(np.uint32), body[:, 1].astype(np.uint32),
            np.array([float(t) for t in body[:, 2]]))


def write_records(path, base, rows, cols, vals):
    with open(path, "w") as f:
        f.write("%d %d\n" % (rows.size, base))
        f.write("".join("%d %d %s\n" % (r, c, repr(float(v)))
                        for r, c, v in zip(rows.tolist(), cols.tolist(), vals.tolist())))


def hostile(base, rows, cols, vals, seed, spread=3, split_every=7):
    """shuffle + duplicate records with order-dependent sums + absent row ids"""
    rng = np.random.default_rng(seed)
    pick = np.arange(rows.size) % split_every == 0
    extra_r, extra_c = rows[pick], cols[pick]
    # v = (v - big) + big' pieces: the folded value depends on the addition order
    big = rng.standard_normal(extra_r.size) * 1e6
    r = np.concatenate([rows, extra_r, extra_r])
    c = np.concatenate([cols, extra_c, extra_c])
    v = np.concatenate([vals, big, -big * (1 + 1e-9)])
    perm = rng.permutation(r.size)
    r, c, v = r[perm], c[perm], v[perm]
    r = ((r.astype(np.int64) - base) * spread + base + 5).astype(np.uint32)
    return r, c, v


def assert_same_csr(got, want):
    nrows, offs, cols, vals = got
    assert nrows == want.nrows
    assert np.array_equal(offs, want.offs)
    assert np.array_equal(cols, want.cols)
    assert vals.tobytes() == want.vals.tobytes()


@pytest.mark.parametrize("name", orc.TOY + orc.NEK)
def test_sorted_files_take_the_fast_path(abi, ctx, name):
    path = orc.matrix_path(name)
    base, r, c, v = read_records(path)
    assert_same_csr(abi.coo_to_csr(ctx, r, c, v), orc.matrix_read(path))


@pytest.mark.parametrize("name", ["A1_02x02", "I1_05x05", "tj7a_A_18", "xn3b_A_10", "xn3b_A_18"])
def test_hostile_records_match_the_readers(abi, ctx, name, tmp_path):
    base, r, c, v = read_records(orc.matrix_path(name))
    r, c, v = hostile(base, r, c, v, seed=len(name))
    f = str(tmp_path / "hostile.txt")
    write_records(f, base, r, c, v)
    want = orc.matrix_read(f)
    got = abi.coo_to_csr(ctx, r, c, v)
    assert_same_csr(got, want)
    ref = orc.ref_matrix_read(f)   # the reference's own reader, when built
    if ref is not None:
        assert_same_csr(got, ref)
    # sizes-only query
    import ctypes as C
    nr, m = C.c_uint32(0), C.c_uint64(0)
    rc = abi.load().b200_coo_to_csr(ctx.h, r.size, r.ctypes.data, c.ctypes.data, v.ctypes.data,
                                    C.byref(nr), C.byref(m), None, None, None)
    assert rc == 0 and nr.value == want.nrows and m.value == want.cols.size


def test_million_random_records(abi, ctx, tmp_path):
    rng = np.random.default_rng(11)
    n = 1_000_000
    r = rng.integers(0, 70_000, n).astype(np.uint32) * 17
    c = rng.integers(0, 50, n).astype(np.uint32)       # many collisions per row
    v = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
    f = str(tmp_path / "rand.txt")
    write_records(f, 0, r, c, v)
    assert_same_csr(abi.coo_to_csr(ctx, r, c, v), orc.matrix_read(f))


def test_edge_cases(abi, ctx):
    one = abi.coo_to_csr(ctx, [7], [9], [2.5])
    assert one[0] == 1 and one[1].tolist() == [0, 1] and one[2].tolist() == [9] and one[3].tolist() == [2.5]
    dup = abi.coo_to_csr(ctx, [3, 3, 3], [4, 4, 4], [1e16, 1.0, -1e16])
    assert dup[0] == 1 and dup[3].tolist() == [(1e16 + 1.0) - 1e16]   # left to right
    # descending input: worst case for the sorted check
    r = np.arange(1000, 0, -1, dtype=np.uint32)
    got = abi.coo_to_csr(ctx, r, r, r.astype(np.float64))
    assert got[0] == 1000 and got[2].tolist() == list(range(1, 1001))
    with pytest.raises(abi.B200Error):
        abi.coo_to_csr(ctx, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0))


@pytest.mark.parametrize("name", ["tj7a_A_18", "xn3b_A_18"])
def test_from_coo_builds_the_cholmod_operator(abi, ctx, name, tmp_path):
    """records -> device layout without a host CSR == the oracle's operator
    (src/cholmod-impl.h:5-21) of the file those records make"""
    base, r, c, v = read_records(orc.matrix_path(name))
    r, c, v = hostile(base, r, c, v, seed=3, spread=1)
    f = str(tmp_path / "h.txt")
    write_records(f, base, r, c, v)
    Mo = orc.op_upper_mirror(orc.matrix_read(f))
    M = abi.Matrix.from_coo(ctx, base, r, c, v, abi.MAT_SYM_UPPER)
    offs, cols, vals = M.export()
    assert np.array_equal(offs, Mo.offs) and np.array_equal(cols, Mo.cols)
    assert vals.tobytes() == Mo.vals.tobytes()
    x = np.random.default_rng(1).standard_normal(Mo.n)
    assert np.array_equal(M.spmv_host(x), orc.spmv_fma(Mo, x))
    M.close()


def test_driver_with_device_ingest(tmp_path):
    """the harness reader with LSBENCH_B200_INGEST=1 solves a shuffled file to
    the same CSV/iteration line as the host reader does"""
    from lsbench_b200 import build_host
    build_host.build()
    base, r, c, v = read_records(orc.matrix_path("xn3b_A_18"))
    perm = np.random.default_rng(5).permutation(r.size)
    f = str(tmp_path / "shuffled.txt")
    write_records(f, base, r[perm], c[perm], v[perm])
    outs = []
    for ing in ("0", "1"):
        env = dict(os.environ, LSBENCH_B200_INGEST=ing)
        p = subprocess.run([build_host.DRIVER, "--solver", "b200", "--matrix", f, "--trials=1"],
                           capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0, p.stderr
        lines = p.stdout.strip().splitlines()
        outs.append(lines[3].split(",")[:5])       # gpus, iterations, status, relres, true_relres
    assert outs[0] == outs[1], outs


# --------------------------------------------------------------------------- text on the device
def body_of(path):
    """(nnz, base, bytes after the header line)"""
    raw = open(path, "rb").read()
    nl = raw.index(b"\n")
    nnz, base = (int(t) for t in raw[:nl].split())
    return nnz, base, raw[nl + 1:]


@pytest.mark.parametrize("name", orc.TOY + orc.NEK)
def test_text_is_parsed_on_the_device(abi, ctx, name):
    """lines found and numbers parsed on the GPU (parse.cuh: strict, exact),
    then sort / fold / compress: the reference's files come out bit for bit"""
    path = orc.matrix_path(name)
    nnz, base, body = body_of(path)
    nr, offs, cols, vals, nhost = abi.text_to_csr(ctx, body, nnz)
    assert_same_csr((nr, offs, cols, vals), orc.matrix_read(path))
    assert nhost <= 1e-4 * nnz + 1          # fixed-point 15-digit values: the exact fast path


def test_text_with_17_digit_values(abi, ctx, tmp_path):
    """shuffled, duplicated records written with 17 significant digits: off
    Clinger's fast path, rounded on the device with exact integer arithmetic
    (parse.cuh b2_exact_decimal); a few records carry values only the host's
    strtod is trusted with.  The result is the readers' CSR, bit for bit"""
    base, r, c, v = read_records(orc.matrix_path("xn3b_A_18"))
    r, c, v = hostile(base, r, c, v, seed=21)
    f = str(tmp_path / "hostile.txt")
    write_records(f, base, r, c, v)
    # values outside what the device takes: tiny / huge exponents, 25 digits, inf
    extra = ["7 7 1.5e-300\n", "7 8 -2.25e+200\n", "8 8 0.1234567890123456789012345\n",
             "9 9 1e-40\n", "9 10 inf\n"]
    raw = open(f).read().split("\n", 1)
    nnz0 = int(raw[0].split()[0])
    with open(f, "w") as g:
        g.write("%d %d\n" % (nnz0 + len(extra), base) + raw[1] + "".join(extra))
    nnz, base2, body = body_of(f)
    nr, offs, cols, vals, nhost = abi.text_to_csr(ctx, body, nnz)
    want = orc.matrix_read(f)
    assert_same_csr((nr, offs, cols, vals), want)
    assert nhost == len(extra)
    ref = orc.ref_matrix_read(f)
    if ref is not None:
        assert_same_csr((nr, offs, cols, vals), ref)


@pytest.mark.parametrize("body,nnz", [
    (b"1 1 2.5\n\n2 2 1.0\n", 2),       # a blank line: fscanf would skip it, a line parser must not guess
    (b" 1 1 2.5\n2 2 1.0\n", 2),        # leading blank
    (b"1 1 2.5 \n2 2 1.0\n", 2),        # trailing blank: the reference requires the newline right there
    (b"1 1 2.5\n2 2\n", 2),             # a field missing
    (b"1 1 2.5\n", 2),                  # fewer lines than records
    (b"1 1 abc\n2 2 1\n", 2),
])
def test_text_that_is_not_one_strict_record_per_line_is_refused(abi, ctx, body, nnz):
    with pytest.raises(abi.B200Error) as e:
        abi.text_to_csr(ctx, body, nnz)
    assert e.value.code == 1      # B200_EINVAL


def test_text_extra_lines_after_the_records_are_ignored(abi, ctx):
    nr, offs, cols, vals, nhost = abi.text_to_csr(ctx, b"3 1 4\n1 2 -0.5\n1 2 1e-3\ngarbage\n", 3)
    assert (nr, offs.tolist(), cols.tolist()) == (2, [0, 1, 2], [2, 1])
    assert vals.tolist() == [-0.5 + 1e-3, 4.0]


def test_harness_reader_fuzz_with_device_ingest(tmp_path, monkeypatch):
    """generated files, strict (one record per line -> parsed on the device) and
    loose (white space the fscanf grammar allows -> refused by the device parser,
    tokenised on the host, sorted / folded on the device): the harness reader with
    LSBENCH_B200_INGEST=1 returns the oracle reader's CSR bit for bit"""
    import ctypes as C
    from lsbench_b200 import build_host
    from test_host_shell import Csr, _random_coo_text
    build_host.build()
    L = C.CDLL(build_host.LIB)
    L.lsbench_matrix_read.restype = C.POINTER(Csr)
    L.lsbench_matrix_read.argtypes = [C.c_char_p]
    L.lsbench_matrix_free.argtypes = [C.POINTER(Csr)]
    monkeypatch.setenv("LSBENCH_B200_INGEST", "1")
    rng = np.random.default_rng(99)
    f = str(tmp_path / "fuzz.txt")
    for k in range(80):
        text = _random_coo_text(rng)
        if k % 2 == 0:   # make it strict: one record per line, single blanks
            head, *recs = [l for l in text.split("\n") if l.strip()]
            text = head + "\n" + "".join(" ".join(r.split()) + "\n" for r in recs)
        with open(f, "w") as g:
            g.write(text)
        want = orc.matrix_read(f)
        p = L.lsbench_matrix_read(f.encode())
        a = p.contents
        nnz = a.offs[a.nrows]
        got = (a.nrows, np.ctypeslib.as_array(a.offs, (a.nrows + 1,)).copy(),
               np.ctypeslib.as_array(a.cols, (nnz,)).copy(), np.ctypeslib.as_array(a.vals, (nnz,)).copy())
        base = int(a.base)
        L.lsbench_matrix_free(p)
        assert base == want.base
        assert_same_csr(got, want)
