"""Small end-to-end pass for `compute-sanitizer --tool memcheck` (one tool per
gpurun call): ingest, layout conversion with the CHOLMOD operator, index
compression, all three SpMV bins, the streaming PCG with and without graphs.
The cluster/DSMEM kernel is left out (B200_PCG_NO_SMALL)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from lsbench_b200 import abi

ctx = abi.Context(0)
A = orc.matrix_read(orc.matrix_path("xn3b_A_18"))
Mo = orc.op_upper_mirror(A)
M = abi.Matrix.from_csr(ctx, A.nrows, A.base, A.offs, A.cols, A.vals, abi.MAT_SYM_UPPER)
offs, cols, vals = M.export()
assert np.array_equal(cols, Mo.cols) and vals.tobytes() == Mo.vals.tobytes()
x = np.random.default_rng(0).standard_normal(Mo.n)
assert np.array_equal(M.spmv_host(x), orc.spmv_fma(Mo, x))
b = orc.rhs(Mo.n)
for fl in (abi.PCG_NO_SMALL, abi.PCG_NO_SMALL | abi.PCG_NO_GRAPH):
    xs, r, rc = M.pcg_host(b, tol=1e-10, maxit=5000, flags=fl)
    assert rc == 0 and r.status == 0 and r.true_relres <= 1e-10
M.close()
# ingest: shuffled records with duplicates
rows = np.repeat(np.arange(A.nrows, dtype=np.uint32), np.diff(A.offs.astype(np.int64))) + A.base
perm = np.random.default_rng(1).permutation(rows.size)
r2 = np.concatenate([rows[perm], rows[:100]])
c2 = np.concatenate([A.cols[perm], A.cols[:100]])
v2 = np.concatenate([A.vals[perm], A.vals[:100]])
nr, o, c, v = abi.coo_to_csr(ctx, r2, c2, v2)
assert nr == A.nrows and np.array_equal(c, A.cols)
# index compression + generators + power-law bins
for kind, size in ((abi.GEN_POISSON7, 96), (abi.GEN_POISSON27, 40), (abi.GEN_POWERLAW, 30000)):
    G = abi.Matrix.generate(ctx, kind, size, seed=3)
    i = G.info()
    y = G.spmv_host(np.ones(i.n_local))
    if kind != abi.GEN_POWERLAW:
        xg, rg, rc = G.pcg_host(orc.rhs(i.n_local), tol=1e-8, flags=abi.PCG_NO_SMALL)
        assert rc == 0 and rg.status == 0
    G.export()
    G.close()
ctx.close()
print("SANITIZE_SMOKE OK")
