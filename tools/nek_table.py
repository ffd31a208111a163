"""BASELINE.json config 2: the Nek coarse-grid matrices on one B200.
b200 PCG (on-chip path and streaming path) next to the CPU direct-solve
stand-in (oracle LDL^T, RCM; CHOLMOD-equivalent restatement, timed as
src/cholmod-impl.h:45-63: factorise once untimed, `trials` solves timed).
   python tools/nek_table.py [trials]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from lsbench_b200 import abi

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
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
    rows.append(out)
    print(json.dumps(out)); sys.stdout.flush()
    M.close()
print("%-10s %6s %8s | %8s %5s | %8s %5s | %9s" % ("matrix", "n", "nnz", "onchip", "its", "stream", "its", "cpu_direct"))
for o in rows:
    print("%-10s %6d %8d | %7.3f  %5d | %7.3f  %5d | %8.3f ms" % (
        o["matrix"], o["n"], o["nnz"], o["onchip"]["ms_per_solve_wall"], o["onchip"]["iters"],
        o["stream"]["ms_per_solve_wall"], o["stream"]["iters"], o["cpu_direct"]["ms_per_solve"]))
