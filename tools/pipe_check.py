"""Round-2 starting point: the software-pipelined fp32-value SpMV kernel
(csrc/sellc32p.cuh) on a GPU -- parity first, then time, beside the plain kernels.

    B200_SPMV_PIPE=1 python tools/pipe_check.py [N]        27-point N^3, default 192
    B200_SPMV_PIPE=2 python tools/pipe_check.py [N]        fp64-stored values pipelined as well
                                                           (128-thread CTAs); compare with the
                                                           "f64" line of a B200_SPMV_PIPE=1 run

Parity: SpMV bits against the oracle's fma product on a 27-point 40^3 and a 7-point
48^3 grid and on ragged rows; PCG iteration count and solution against the fp64-stored
matrix.  Time: b200_spmv_time and the per-kernel PCG times.  The kernel's indexing is
already checked on the CPU (tests/test_spmv_emul.py); it has not run on hardware yet.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
from lsbench_b200 import abi  # noqa: E402

if os.environ.get("B200_SPMV_PIPE", "0") in ("", "0"):
    raise SystemExit("set B200_SPMV_PIPE=1 (the switch is read once per process)")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 192
abi.load()
ctx = abi.Context(0)
rng = np.random.default_rng(0)


def csr(M):
    return M.n, 0, M.offs.astype(np.uint32), M.cols, M.vals


for M in (orc.gen_poisson27(40), orc.gen_poisson7(48)):
    Md = abi.Matrix.from_csr(ctx, *csr(M), abi.MAT_VALUES_F32)
    assert Md.info().values_f32 == 1
    x = rng.standard_normal(M.n)
    assert Md.spmv_host(x).tobytes() == orc.spmv_fma(M, x).tobytes(), "stencil SpMV bits differ"
    b = orc.rhs(M.n)
    xs, r, rc = Md.pcg_host(b, flags=abi.PCG_NO_SMALL)
    assert rc == 0 and r.status == 0 and orc.true_relres(M, b, xs) <= 1e-10
    Md.close()
n = 4000
lens = (np.arange(n) % 32) + 1
offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
cols = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]).astype(np.uint32)
M = orc.Op(n, offs, cols, rng.integers(-1000, 1000, int(offs[-1])).astype(np.float64) / 64.0)
for fl in (abi.MAT_VALUES_F32, abi.MAT_VALUES_F32 | abi.MAT_NO_SORT):
    Md = abi.Matrix.from_csr(ctx, *csr(M), fl)
    x = rng.standard_normal(n)
    assert Md.spmv_host(x).tobytes() == orc.spmv_fma(M, x).tobytes(), "ragged SpMV bits differ"
    Md.close()
print("parity: ok")

out = {"workload": "poisson27:%d" % N, "B200_SPMV_PIPE": os.environ.get("B200_SPMV_PIPE")}
LEVEL = int(os.environ["B200_SPMV_PIPE"])
if LEVEL >= 2:   # fp64-stored values through the pipelined kernel: bits first
    for M in (orc.gen_poisson27(40), orc.gen_poisson7(48)):
        Md = abi.Matrix.from_csr(ctx, *csr(M), 0)
        x = rng.standard_normal(M.n)
        assert Md.spmv_host(x).tobytes() == orc.spmv_fma(M, x).tobytes(), "fp64 pipelined SpMV bits differ"
        Md.close()
    print("parity (fp64 values, pipelined): ok")
for mname, mflags in (("f64_pipelined" if LEVEL >= 2 else "f64", 0), ("f32_pipelined", abi.MAT_VALUES_F32)):
    Md = abi.Matrix.generate(ctx, abi.GEN_POISSON27, N, 1, mflags)
    i = Md.info()
    n = i.n_local
    dx, dy = abi.DeviceArray(ctx, n), abi.DeviceArray(ctx, n)
    dx.upload(rng.standard_normal(n))
    ms = min(Md.spmv_time(dx, dy, reps=30) for _ in range(3))
    b = np.arange(n, dtype=np.float64)
    Md.pcg_host(b, flags=abi.PCG_NO_SMALL)
    xs, r, _ = Md.pcg_host(b, flags=abi.PCG_NO_SMALL | abi.PCG_TIME_KERNELS)
    out[mname] = {"spmv_ms": ms, "stored_gbs": (i.matrix_stream_bytes + 16 * n) / ms / 1e6,
                  "iters": r.iters, "ms_per_it": r.solve_ms / max(r.iters, 1),
                  "kernel_ms": [r.spmv_ms, r.update_ms, r.pupdate_ms], "true_relres": r.true_relres}
    Md.close()
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pipe_check.json"), "w"), indent=1)
