"""The on-chip coarse-grid kernel (csrc/small.cu) on the seven Nek matrices: Jacobi and
Chebyshev-Jacobi of degree 2 / 3 -- iterations, kernel ms, wall ms per solve, residual,
distance to the direct solve (tests/golden/direct.npz).
    python tools/small_check.py [trials]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
from lsbench_b200 import abi  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 50
gold = np.load(os.path.join(ROOT, "tests", "golden", "direct.npz"))
ctx = abi.Context(0)
rows = []
for name in orc.NEK:
    A = orc.matrix_read(orc.matrix_path(name))
    Mo = orc.op_upper_mirror(A)
    b = orc.rhs(Mo.n)
    M = abi.Matrix.from_csr(ctx, A.nrows, A.base, A.offs, A.cols, A.vals, abi.MAT_SYM_UPPER)
    db, dx = ctx.array(Mo.n).upload(b), ctx.array(Mo.n)
    out = {"matrix": name, "n": int(Mo.n)}
    modes = (("jacobi", 0), ("bj", abi.PCG_BLOCK_JACOBI), ("cheb2", abi.PCG_CHEBYSHEV2), ("cheb3", abi.PCG_CHEBYSHEV3))
    if os.environ.get("SMALL_CHECK_MODES"):
        modes = tuple(m for m in modes if m[0] in os.environ["SMALL_CHECK_MODES"].split(","))
    for label, fl in modes:
        try:
            for _ in range(3):
                dx.zero()
                M.pcg(db, dx, flags=fl)
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(trials):
                dx.zero()
                r, rc = M.pcg(db, dx, flags=fl)
            ctx.sync()
            dt = (time.perf_counter() - t0) / trials
            x = dx.download()
            out[label] = {"ms_wall": round(dt * 1e3, 4), "ms_kernel": round(r.solve_ms, 4), "iters": r.iters,
                          "status": r.status, "degree": r.outer_iters, "path": r.path,
                          "block_jacobi": r.block_jacobi,
                          "true_relres": r.true_relres, "replacements": r.replacements,
                          "rel_diff_direct": float(np.linalg.norm(x - gold[name]) / np.linalg.norm(gold[name]))}
        except abi.B200Error as e:
            out[label] = {"error": str(e)[:200]}
    rows.append(out)
    print(json.dumps(out), flush=True)
    M.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "small_check.json"), "w"), indent=1)
keys = [m[0] for m in modes]
print("%-10s | " % "matrix" + " | ".join("%-22s" % (k + " its / ms") for k in keys))
for o in rows:
    print("%-10s | " % o["matrix"] + " | ".join(
        "%5d / %7.3f ms      " % (o[k]["iters"], o[k]["ms_wall"]) if "iters" in o.get(k, {}) else "error                 "
        for k in keys))
