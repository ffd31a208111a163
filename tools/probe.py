"""Quick device probe: SpMV GB/s and PCG timing on a generated grid.
usage: python tools/probe.py [poisson7|poisson27|powerlaw] [size] [flags]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lsbench_b200 import abi

kind = sys.argv[1] if len(sys.argv) > 1 else "poisson7"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mode = sys.argv[4] if len(sys.argv) > 4 else "all"   # all | pcg (short solve only, for ncu)
k = {"poisson7": 1, "poisson27": 2, "powerlaw": 3}[kind]
ctx = abi.Context(0)
t = time.time()
M = abi.Matrix.generate(ctx, k, size, seed=1, flags=flags)
i = M.info()
print("generate+convert %.2fs n=%d nnz=%d padded=%d sell_rows=%d vec=%d long=%d sigma=%d perm=%d wmax=%d dev=%.2fGB "
      "uniform_slices=%d/%d stream=%.3f B/nnz" % (
    time.time() - t, i.n_local, i.nnz, i.nnz_padded, i.sell_rows, i.vec_rows, i.long_rows,
    i.sell_sigma, i.sell_perm, i.sell_max_width, i.device_bytes / 1e9,
    i.sell_uniform_slices, i.sell_slices, i.matrix_stream_bytes / max(i.nnz, 1)))
n = i.n_local
sp, it = M.algorithmic_bytes()
dx, dy = ctx.array(n), ctx.array(n)
dx.upload(np.random.default_rng(0).standard_normal(n))
for rep in range(3 if mode == "all" else 0):
    ms = M.spmv_time(dx, dy, reps=20)
    print("spmv %.4f ms  %.1f GB/s (algorithmic %d B)" % (ms, sp / ms / 1e6, sp))
if kind != "powerlaw":
    b = np.arange(n, dtype=np.float64)
    db, dxx = ctx.array(n).upload(b), ctx.array(n)
    for fl in ((0, abi.PCG_TIME_KERNELS) if mode == "all" else (abi.PCG_NO_GRAPH,)):
        dxx.zero()
        r, rc = M.pcg(db, dxx, tol=1e-10, maxit=20000 if mode == "all" else 12, flags=fl | abi.PCG_NO_SMALL)
        print("pcg flags=%d iters=%d status=%d relres=%.3e true=%.3e solve=%.2f ms  %.4f ms/it  %.1f GB/s  classes spmv=%.4f upd=%.4f pupd=%.4f" % (
            fl, r.iters, r.status, r.relres, r.true_relres, r.solve_ms, r.solve_ms / max(r.iters, 1),
            it * r.iters / r.solve_ms / 1e6, r.spmv_ms, r.update_ms, r.pupdate_ms))
M.close(); ctx.close()
