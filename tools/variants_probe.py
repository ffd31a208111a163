"""Per-kernel times of the opt-in variants beside the default path (one GPU):
fp32-stored values (B200_MAT_VALUES_F32) and single-reduction CG.

    python tools/variants_probe.py [N]      27-point N^3, default 192
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsbench_b200 import abi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 192
abi.load()
ctx = abi.Context(0)
out = {"workload": "poisson27:%d" % N}
for mname, mflags in (("f64", 0), ("f32", abi.MAT_VALUES_F32)):
    M = abi.Matrix.generate(ctx, abi.GEN_POISSON27, N, 1, mflags)
    i = M.info()
    n = i.n_local
    b = np.arange(n, dtype=np.float64)
    dx, dy = abi.DeviceArray(ctx, n), abi.DeviceArray(ctx, n)
    dx.upload(np.random.default_rng(0).standard_normal(n))
    ms = min(M.spmv_time(dx, dy, reps=30) for _ in range(3))
    sb, _ = M.algorithmic_bytes()
    out[mname] = {"values_f32": i.values_f32, "stream_bytes": i.matrix_stream_bytes,
                  "spmv_ms": ms, "spmv_alg_gbs": sb / ms / 1e6,
                  "spmv_stored_gbs": (i.matrix_stream_bytes + 16 * n) / ms / 1e6}
    for pname, pflags in (("pcg", 0), ("pcg_sr", abi.PCG_SINGLE_REDUCTION)):
        fl = abi.PCG_NO_SMALL | pflags
        M.pcg_host(b, flags=fl)                                  # warm (graph capture)
        x, r, _ = M.pcg_host(b, flags=fl)
        xt, rt, _ = M.pcg_host(b, flags=fl | abi.PCG_TIME_KERNELS)
        out[mname][pname] = {"iters": r.iters, "status": r.status, "solve_ms": r.solve_ms,
                             "ms_per_it": r.solve_ms / max(r.iters, 1), "true_relres": r.true_relres,
                             "kernel_ms": [rt.spmv_ms, rt.update_ms, rt.pupdate_ms]}
    M.close()
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r01n_variants_probe.json", "w"), indent=1)
