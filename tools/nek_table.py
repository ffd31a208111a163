"""BASELINE.json config 2: the Nek coarse-grid matrices on one B200.
b200 PCG (on-chip path and streaming path) next to the CPU direct-solve
stand-in (oracle LDL^T, RCM; CHOLMOD-equivalent restatement, timed as
src/cholmod-impl.h:45-63: factorise once untimed, `trials` solves timed) and
the reference's own GPU backend, UNMODIFIED (src/cusparse.c = cuSOLVER-Sp
sparse Cholesky, compiled from the reference sources into
oracle/_ref/driver_cusolver by `make -C oracle ref-cusolver`; it refactors in
every timed call, src/cusparse.c:189-197, default ordering RCM).
   python tools/nek_table.py [trials] [cusolver_trials]
"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from lsbench_b200 import abi

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cu_trials = int(sys.argv[2]) if len(sys.argv) > 2 else 5
CU_DRIVER = os.path.join(ROOT, "oracle", "_ref", "driver_cusolver")


def cusolver_ms(path):
    """elapsed column of the reference's CSV row (total clock() seconds over
    `trials` solves, src/cusparse.c:207-209) -> ms per solve."""
    if not os.path.exists(CU_DRIVER) or cu_trials < 1:
        return None
    try:
        r = subprocess.run([CU_DRIVER, "--solver", "cusolver", "--matrix", path,
                            "--trials=%d" % cu_trials], capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        return {"error": "timeout"}
    for line in r.stdout.splitlines():
        f = line.strip().split(",")
        if len(f) == 7 and not line.startswith("==="):
            try:
                return {"ms_per_solve": float(f[6]) / cu_trials * 1e3, "trials": cu_trials,
                        "kind": "reference src/cusparse.c (cusolverSpDcsrlsvchol, RCM), clock()"}
            except ValueError:
                pass
    return {"error": (r.stderr or r.stdout)[-200:]}

gold = np.load(os.path.join(ROOT, "tests", "golden", "direct.npz"))
ctx = abi.Context(0)
rows = []
for name in orc.NEK:
    A = orc.matrix_read(orc.matrix_path(name))
    Mo = orc.op_upper_mirror(A)
    b = orc.rhs(Mo.n)
    M = abi.Matrix.from_csr(ctx, A.nrows, A.base, A.offs, A.cols, A.vals, abi.MAT_SYM_UPPER)
    db, dx = ctx.array(Mo.n).upload(b), ctx.array(Mo.n)
    out = {"matrix": name, "n": Mo.n, "nnz": Mo.nnz}
    for label, fl in (("onchip", 0), ("stream", abi.PCG_NO_SMALL)):
        for _ in range(5):
            dx.zero(); M.pcg(db, dx, flags=fl)
        ctx.sync(); t0 = time.perf_counter()
        for _ in range(trials):
            dx.zero(); r, rc = M.pcg(db, dx, flags=fl)
        ctx.sync(); dt = (time.perf_counter() - t0) / trials
        x = dx.download()
        out[label] = {"ms_per_solve_wall": dt * 1e3, "ms_kernel": r.solve_ms, "iters": r.iters,
                      "true_relres": r.true_relres, "path": r.path,
                      "rel_diff_direct": float(np.linalg.norm(x - gold[name]) / np.linalg.norm(gold[name]))}
    t0 = time.perf_counter(); F = orc.Ldlt(Mo, 1); tf = time.perf_counter() - t0
    F.solve(b); t0 = time.perf_counter()
    for _ in range(trials):
        xd = F.solve(b)
    ts = (time.perf_counter() - t0) / trials
    out["cpu_direct"] = {"factor_s": tf, "ms_per_solve": ts * 1e3, "lnz": int(F.nnz), "cores": 1,
                         "kind": "CHOLMOD-equivalent restatement (oracle LDL^T, RCM)"}
    out["ref_cusolver"] = cusolver_ms(orc.matrix_path(name))
    rows.append(out)
    print(json.dumps(out)); sys.stdout.flush()
    M.close()
print("%-10s %6s %8s | %8s %5s | %8s %5s | %9s | %12s" % ("matrix", "n", "nnz", "onchip", "its", "stream", "its", "cpu_direct", "ref_cusolver"))
for o in rows:
    cu = o.get("ref_cusolver") or {}
    print("%-10s %6d %8d | %7.3f  %5d | %7.3f  %5d | %8.3f ms | %9s ms" % (
        o["matrix"], o["n"], o["nnz"], o["onchip"]["ms_per_solve_wall"], o["onchip"]["iters"],
        o["stream"]["ms_per_solve_wall"], o["stream"]["iters"], o["cpu_direct"]["ms_per_solve"],
        "%.3f" % cu["ms_per_solve"] if "ms_per_solve" in cu else "n/a"))
