"""The C host path on a GPU: `driver --solver b200` (host C -> C ABI -> CUDA),
i.e. the call a user of the reference makes, against the golden direct solve.
Also runs the reference's own, unmodified bin/driver.c when it was built."""
import os
import subprocess

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIRECT = np.load(os.path.join(ROOT, "tests", "golden", "direct.npz"))


@pytest.fixture(scope="module")
def bh():
    from lsbench_b200 import build, build_host
    build.build()
    build_host.build()
    return build_host


def parse(stdout):
    lines = stdout.strip().splitlines()
    i = lines.index("===matrix,n,nnz,trials,solver,ordering,elapsed===")
    row = lines[i + 1].split(",")
    j = [k for k, l in enumerate(lines) if l.startswith("===b200:")][0]
    ext = lines[j + 1].split(",")
    return row, ext


@pytest.mark.parametrize("name", ["I1_05x05", "tj7a_A_12", "xn3b_A_10", "xn3b_A_18"])
def test_driver_b200_matches_direct_solve(bh, name, tmp_path):
    A = orc.matrix_read(orc.matrix_path(name))
    out = str(tmp_path / "x.bin")
    r = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", orc.matrix_path(name),
                        "--trials=3", "--verbose=1", "--dump-x", out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    row, ext = parse(r.stdout)
    # matrix,n,nnz,trials,solver,ordering,elapsed  (src/cusparse.c:207-209)
    assert row[1:6] == [str(A.nrows), str(A.nnz), "3", "6", "0"] and float(row[6]) > 0
    gpus, iters, status, relres, true_relres = int(ext[0]), int(ext[1]), int(ext[2]), float(ext[3]), float(ext[4])
    assert (gpus, status) == (1, 0) and relres <= 1e-10 and true_relres <= 1e-10
    x = np.fromfile(out)
    g = DIRECT[name]
    assert np.linalg.norm(x - g) / np.linalg.norm(g) <= 1e-8
    M = orc.op_upper_mirror(A)
    assert orc.true_relres(M, orc.rhs(M.n), x) <= 1e-10


def test_unmodified_reference_driver_runs_b200(bh):
    if not os.path.exists(bh.DRIVER_REF):
        pytest.skip("driver_ref was not built (no reference tree at build time)")
    r = subprocess.run([bh.DRIVER_REF, "--solver", "b200", "--matrix",
                        orc.matrix_path("tj7a_A_18"), "--trials=2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    row, ext = parse(r.stdout)
    assert row[1] == "3707" and int(ext[2]) == 0 and float(ext[4]) <= 1e-10


def test_driver_full_operator_switch(bh, tmp_path):
    """LSBENCH_B200_OPERATOR=full solves the matrix as stored (what cuSOLVER is
    handed): a different answer, by ~1e-7."""
    name = "tj7a_A_18"
    out = str(tmp_path / "x.bin")
    env = dict(os.environ, LSBENCH_B200_OPERATOR="full")
    r = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", orc.matrix_path(name),
                        "--trials=1", "--dump-x", out], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    x, g = np.fromfile(out), DIRECT[name]
    d = np.linalg.norm(x - g) / np.linalg.norm(g)
    assert 1e-8 < d < 1e-5
    A = orc.matrix_read(orc.matrix_path(name))
    assert orc.true_relres(orc.op_full(A), orc.rhs(A.nrows), x) <= 1e-10


def test_driver_synthetic_and_multi_gpu(bh, tmp_path):
    from lsbench_b200 import abi
    out = str(tmp_path / "x.bin")
    r = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", "poisson27:40",
                        "--trials=2", "--dump-x", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    row, ext = parse(r.stdout)
    M = orc.gen_poisson27(40)
    assert row[1:3] == [str(M.n), str(M.nnz)] and int(ext[2]) == 0
    x1 = np.fromfile(out)
    assert orc.true_relres(M, orc.rhs(M.n), x1) <= 1e-10
    if abi.device_count() >= 2:
        env = dict(os.environ, LSBENCH_B200_NGPUS="2")
        r = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", "poisson27:40",
                            "--trials=2", "--dump-x", out], capture_output=True, text=True,
                           env=env, timeout=600)
        assert r.returncode == 0, r.stderr
        row2, ext2 = parse(r.stdout)
        assert int(ext2[0]) == 2 and int(ext2[2]) == 0 and row2[2] == row[2]
        x2 = np.fromfile(out)
        assert np.linalg.norm(x2 - x1) / np.linalg.norm(x1) <= 1e-9


@pytest.mark.parametrize("name,env", [("xn3b_A_10", {"LSBENCH_B200_ORDERING": "rcm"}),
                                      ("xn3b_A_18", {"LSBENCH_B200_ORDERING": "cli"}),
                                      ("tj7a_A_18", {"LSBENCH_B200_ORDERING": "rcm", "LSBENCH_B200_OPERATOR": "full"})])
def test_driver_applies_the_ordering(bh, name, env, tmp_path):
    """with LSBENCH_B200_ORDERING the solve runs on P A P^T (RCM
    computed on the host, src/cusparse.c:66-85 being the reference's precedent) and
    x comes back in the caller's numbering: same answer as without, to the bar"""
    A = orc.matrix_read(orc.matrix_path(name))
    out = str(tmp_path / "x.bin")
    r = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", orc.matrix_path(name),
                        "--trials=2", "--verbose=1", "--ordering=RCM", "--dump-x", out],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stderr
    row, ext = parse(r.stdout)
    assert row[1:6] == [str(A.nrows), str(A.nnz), "2", "6", "0"]
    assert int(ext[2]) == 0 and float(ext[4]) <= 1e-10
    assert "b200: ordering=rcm bandwidth" in r.stdout
    x = np.fromfile(out)
    full = env.get("LSBENCH_B200_OPERATOR") == "full"
    M = orc.op_full(A) if full else orc.op_upper_mirror(A)
    assert orc.true_relres(M, orc.rhs(M.n), x) <= 1e-10
    if not full:
        g = DIRECT[name]
        assert np.linalg.norm(x - g) / np.linalg.norm(g) <= 1e-8


def test_driver_precision_fp32_and_chebyshev(bh, tmp_path):
    """--precision FP32 (fp32-stored operator: rounded on a Nek file -> refinement
    passes, lossless on a stencil) and LSBENCH_B200_PCG=cheb2 (Chebyshev-Jacobi on the
    on-chip kernel) through the harness: the same bars as the default run"""
    name = "xn3b_A_18"
    A = orc.matrix_read(orc.matrix_path(name))
    out = str(tmp_path / "x.bin")
    r = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", orc.matrix_path(name), "--trials=2",
                        "--precision=FP32", "--verbose=1", "--dump-x", out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    row, ext = parse(r.stdout)
    assert int(ext[2]) == 0 and float(ext[4]) <= 1e-10
    assert "precision=fp32 values rounded" in r.stdout
    x, g = np.fromfile(out), DIRECT[name]
    assert np.linalg.norm(x - g) / np.linalg.norm(g) <= 1e-8
    r = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", "poisson27:40", "--trials=2",
                        "--precision=FP32", "--verbose=1", "--dump-x", out],
                       capture_output=True, text=True, timeout=600,
                       )
    assert r.returncode == 0, r.stderr
    row, ext = parse(r.stdout)
    assert int(ext[2]) == 0 and "precision=fp32 values lossless" in r.stdout
    M = orc.gen_poisson27(40)
    assert orc.true_relres(M, orc.rhs(M.n), np.fromfile(out)) <= 1e-10
    r0 = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", orc.matrix_path(name), "--trials=2"],
                        capture_output=True, text=True, timeout=600,
                        env=dict(os.environ, LSBENCH_B200_PCG="jacobi"))
    r = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", orc.matrix_path(name), "--trials=2",
                        "--dump-x", out], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, LSBENCH_B200_PCG="cheb2"))
    assert r.returncode == 0 and r0.returncode == 0, r.stderr
    (_, e0), (_, e2) = parse(r0.stdout), parse(r.stdout)
    assert int(e2[2]) == 0 and float(e2[4]) <= 1e-10 and int(e2[7]) == 1
    assert int(e0[8]) == 0 and int(e0[7]) == 1
    assert int(e2[1]) < 0.62 * int(e0[1])                      # about half of Jacobi's iterations
    x = np.fromfile(out)
    assert np.linalg.norm(x - g) / np.linalg.norm(g) <= 1e-8
    # the default on the on-chip path: block-Jacobi (fewer iterations than Jacobi, same bars)
    rb = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", orc.matrix_path(name), "--trials=2",
                         "--dump-x", out], capture_output=True, text=True, timeout=600)
    assert rb.returncode == 0, rb.stderr
    _, eb = parse(rb.stdout)
    assert int(eb[2]) == 0 and float(eb[4]) <= 1e-10 and int(eb[7]) == 1 and int(eb[8]) in (16, 32)
    assert int(eb[1]) < 0.9 * int(e0[1])
    x = np.fromfile(out)
    assert np.linalg.norm(x - g) / np.linalg.norm(g) <= 1e-8
    bad = subprocess.run([bh.DRIVER, "--solver", "b200", "--matrix", orc.matrix_path(name), "--trials=1"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, LSBENCH_B200_PCG="ilu"))
    assert bad.returncode != 0 and "LSBENCH_B200_PCG" in bad.stderr


def test_stock_lsbench_tree_with_b200_dropped_in(bh):
    """oracle/_ref/driver_ref_b200: the REFERENCE's lsbench.c / lsbench-csr.c /
    bin/driver.c with the five registration edits and this repository's b200.c
    (built by oracle/dropin.py where the reference tree is mounted).  The stock
    harness runs the b200 backend on a B200: the reference's CSV row, convergence
    to 1e-10, and the same iteration count as this tree's own driver."""
    drop = os.path.join(ROOT, "oracle", "_ref", "driver_ref_b200")
    if not os.path.exists(drop):
        pytest.skip("oracle/_ref/driver_ref_b200 was not built (no reference tree at build time)")
    name = "tj7a_A_18"
    A = orc.matrix_read(orc.matrix_path(name))
    runs = []
    for exe in (drop, bh.DRIVER):
        r = subprocess.run([exe, "--solver", "b200", "--matrix", orc.matrix_path(name), "--trials=2"],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        runs.append(parse(r.stdout))
    (row, ext), (row0, ext0) = runs
    assert row[1:6] == [str(A.nrows), str(A.nnz), "2", "6", "0"] == row0[1:6]
    assert int(ext[2]) == 0 and float(ext[4]) <= 1e-10
    assert ext[1] == ext0[1] and ext[3] == ext0[3]      # iterations and residual: the same solve
