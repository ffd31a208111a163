"""Per-phase clock64 profile of the on-chip PCG kernel (csrc/small.cu):
   B200_SMALL_PROFILE=1 python tools/smallprof.py
prints cycles per iteration for SpMV, the two all-reduce waits, the update and
the window update, as seen by thread 0 of CTA 0."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, orc
from lsbench_b200 import abi
ctx = abi.Context(0)
for name in ("tj7a_A_12", "xn3b_A_10"):
    A = orc.matrix_read(orc.matrix_path(name))
    M = abi.Matrix.from_csr(ctx, A.nrows, A.base, A.offs, A.cols, A.vals, abi.MAT_SYM_UPPER)
    b = orc.rhs(A.nrows)
    for _ in range(2):
        x, r, rc = M.pcg_host(b, tol=1e-10, maxit=5000)
    print(name, r.iters, r.solve_ms)
    M.close()
