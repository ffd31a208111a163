"""The bulk-copy-fed SELL kernel (csrc/sell_tma.cuh) on a GPU: parity first, then time
beside the plain kernel for several ring shapes.

    python tools/tma_check.py [N]                 27-point N^3, default 256

Parity: SpMV bits against the oracle's fma product on stencils and ragged rows (uniform
and explicit slices), PCG iteration count and x bit for bit against the plain kernel.
Time: b200_spmv_time and the per-kernel PCG times, one process per configuration (the
switches are read once per process).
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def csr(M):
    return M.n, 0, M.offs.astype(np.uint32), M.cols, M.vals


if len(sys.argv) > 1 and sys.argv[1] == "--parity":
    import orc
    from lsbench_b200 import abi
    abi.load()
    ctx = abi.Context(0)
    rng = np.random.default_rng(0)
    out = {}
    for name, M in (("poisson27:40", orc.gen_poisson27(40)), ("poisson7:48", orc.gen_poisson7(48))):
        for fl in (0, abi.MAT_VALUES_F32):
            Md = abi.Matrix.from_csr(ctx, *csr(M), fl)
            x = rng.standard_normal(M.n)
            assert Md.spmv_host(x).tobytes() == orc.spmv_fma(M, x).tobytes(), "stencil SpMV bits differ"
            b = orc.rhs(M.n)
            xs, r, rc = Md.pcg_host(b, flags=abi.PCG_NO_SMALL)
            assert rc == 0 and r.status == 0 and orc.true_relres(M, b, xs) <= 1e-10
            out["%s/%d" % (name, fl)] = [r.iters, float(np.linalg.norm(xs)), xs.tobytes().hex()[:32]]
            Md.close()
    n = 4000
    lens = (np.arange(n) % 32) + 1
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    cols = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]).astype(np.uint32)
    M = orc.Op(n, offs, cols, rng.integers(-1000, 1000, int(offs[-1])).astype(np.float64) / 64.0)
    for fl in (0, abi.MAT_NO_SORT, abi.MAT_VALUES_F32):
        Md = abi.Matrix.from_csr(ctx, *csr(M), fl)
        x = rng.standard_normal(n)
        assert Md.spmv_host(x).tobytes() == orc.spmv_fma(M, x).tobytes(), "ragged SpMV bits differ"
        Md.close()
    for name in ("tj7a_A_18", "xn3b_A_10"):
        A = orc.matrix_read(orc.matrix_path(name))
        Mo = orc.op_upper_mirror(A)
        Md = abi.Matrix.from_csr(ctx, A.nrows, A.base, A.offs, A.cols, A.vals, abi.MAT_SYM_UPPER)
        x = rng.standard_normal(Mo.n)
        assert Md.spmv_host(x).tobytes() == orc.spmv_fma(Mo, x).tobytes(), "Nek SpMV bits differ"
        xs, r, rc = Md.pcg_host(orc.rhs(Mo.n), flags=abi.PCG_NO_SMALL)
        out[name] = [r.iters, float(np.linalg.norm(xs)), xs.tobytes().hex()[:32]]
        Md.close()
    print(json.dumps(out))
    sys.exit(0)

if len(sys.argv) > 1 and sys.argv[1] == "--time":
    from lsbench_b200 import abi
    N = int(sys.argv[2])
    abi.load()
    ctx = abi.Context(0)
    res = {}
    for mname, mflags in (("f64", 0), ("f32", abi.MAT_VALUES_F32)):
        Md = abi.Matrix.generate(ctx, abi.GEN_POISSON27, N, 1, mflags)
        i = Md.info()
        n = i.n_local
        dx, dy = abi.DeviceArray(ctx, n), abi.DeviceArray(ctx, n)
        dx.upload(np.random.default_rng(0).standard_normal(n))
        ms = min(Md.spmv_time(dx, dy, reps=30) for _ in range(3))
        b = np.arange(n, dtype=np.float64)
        Md.pcg_host(b, flags=abi.PCG_NO_SMALL)
        xs, r, _ = Md.pcg_host(b, flags=abi.PCG_NO_SMALL | abi.PCG_TIME_KERNELS)
        xg, rg, _ = Md.pcg_host(b, flags=abi.PCG_NO_SMALL)
        res[mname] = {"spmv_ms": round(ms, 4), "stored_gbs": round((i.matrix_stream_bytes + 16 * n) / ms / 1e6, 1),
                      "iters": rg.iters, "ms_per_it": round(rg.solve_ms / max(rg.iters, 1), 4),
                      "kernel_ms": [round(v, 4) for v in (r.spmv_ms, r.update_ms, r.pupdate_ms)],
                      "true_relres": rg.true_relres, "replacements": rg.replacements}
        Md.close()
        dx.free(), dy.free()
    print(json.dumps(res))
    sys.exit(0)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def sub(args, **env):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, __file__] + args, env=e, capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        return {"error": (r.stderr or r.stdout)[-400:]}
    return json.loads(r.stdout.strip().splitlines()[-1])


plain = sub(["--parity"], B200_SPMV_TMA=0)
tma = sub(["--parity"], B200_SPMV_TMA=1)
tma12 = sub(["--parity"], B200_SPMV_TMA=1, B200_TMA_WARPS=12, B200_TMA_SMEM_KB=168)
ok = "error" not in plain and plain == tma == tma12
print("parity: %s" % ("ok (SpMV bits = oracle fma; PCG iterations and x identical to the plain kernel)"
                      if ok else "FAILED"), flush=True)
if not ok:
    print(json.dumps({"plain": plain, "tma": tma, "tma12": tma12})[:3000])
    sys.exit(1)
out = []
for cfg in ({"B200_SPMV_TMA": 0},
            {"B200_SPMV_TMA": 1, "B200_TMA_WARPS": 8, "B200_TMA_SMEM_KB": 168},
            {"B200_SPMV_TMA": 1, "B200_TMA_WARPS": 12, "B200_TMA_SMEM_KB": 168},
            {"B200_SPMV_TMA": 1, "B200_TMA_WARPS": 12, "B200_TMA_SMEM_KB": 200}):
    r = sub(["--time", str(N)], **cfg)
    line = {"workload": "poisson27:%d" % N, "cfg": cfg, "res": r}
    print(json.dumps(line), flush=True)
    out.append(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tma_check.json"), "w"), indent=1)
