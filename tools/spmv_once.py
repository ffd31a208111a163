"""A few SpMV launches on a generated stencil operator, for profiler captures:
    python tools/spmv_once.py [poisson27|poisson7] [N] [f32]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsbench_b200 import abi  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "poisson27"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 192
fl = abi.MAT_VALUES_F32 if len(sys.argv) > 3 and sys.argv[3] == "f32" else 0
ctx = abi.Context(0)
M = abi.Matrix.generate(ctx, abi.GEN_POISSON27 if kind == "poisson27" else abi.GEN_POISSON7, N, 1, fl)
n = M.info().n_local
dx, dy = abi.DeviceArray(ctx, n), abi.DeviceArray(ctx, n)
dx.upload(np.random.default_rng(0).standard_normal(n))
print("ms per spmv", min(M.spmv_time(dx, dy, reps=10) for _ in range(2)))
