"""Round-2 starting point for B200_MAT_COL_BLOCK (column-blocked SpMV for the power-law
operator, BASELINE.json config 5): parity on the GPU first, then time against the
unblocked layout for several block widths.  Written after round 1's GPU budget was
spent; the kernels are checked on the host emulator (tests/test_pcg_emul.py), the
host orchestration (csrc/convert.cu build_layout_or_blocks) has not run yet.

    python tools/colblock_check.py [rows]          default 50 000 000
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) > 2 and sys.argv[1] == "--time":
    # one block width per process: B200_COL_BLOCK_MB is read at conversion time
    from lsbench_b200 import abi
    rows, blocked = int(sys.argv[2]), int(sys.argv[3])
    abi.load()
    ctx = abi.Context(0)
    M = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, rows, seed=1, flags=abi.MAT_COL_BLOCK if blocked else 0)
    i = M.info()
    n = i.n_local
    dx, dy = ctx.array(n), ctx.array(n)
    dx.upload(np.random.default_rng(0).standard_normal(n))
    ms = min(M.spmv_time(dx, dy, reps=10) for _ in range(3))
    sp, _ = M.algorithmic_bytes()
    print(json.dumps({"rows": rows, "col_block_mb": os.environ.get("B200_COL_BLOCK_MB") if blocked else None,
                      "kernel": os.environ.get("B200_COL_BLOCK_KERNEL", "grouped") if blocked else None,
                      "sigma": os.environ.get("B200_COL_BLOCK_SIGMA", "131072") if blocked else None,
                      "col_blocks": i.col_blocks, "ms_per_spmv": ms, "algorithmic_gbs": sp / ms / 1e6,
                      "matrix_stream_bytes": i.matrix_stream_bytes, "device_GB": i.device_bytes / 1e9}))
    sys.exit(0)

import orc  # noqa: E402
from lsbench_b200 import abi  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
abi.load()
ctx = abi.Context(0)
# ---- parity: 1.7 M rows, 1 MB of x per range = 13 column ranges ----------------------------------
os.environ["B200_COL_BLOCK_MB"] = "1"
n = 1_700_000
Mo = orc.gen_powerlaw(n, 1, 0, 4096)     # the oracle's first 4096 rows, global columns
M = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, n, seed=1, flags=abi.MAT_COL_BLOCK)
i = M.info()
assert i.col_blocks == 13, i.col_blocks
x = np.random.default_rng(0).standard_normal(n)
y = M.spmv_host(x)
ref, scale = orc.spmv(Mo, x, want_abs=True)
assert np.all(np.abs(y[:4096] - ref) <= 1e-13 * np.maximum(scale, 1e-300)), "blocked SpMV differs"
M0 = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, n, seed=1, flags=0)
y0 = M0.spmv_host(x)
assert np.max(np.abs(y - y0)) <= 1e-11 * np.max(np.abs(y0)), "blocked vs unblocked"
o0, c0, v0 = M0.export()
o1, c1, v1 = M.export()
assert np.array_equal(o0, o1) and np.array_equal(c0, c1) and v0.tobytes() == v1.tobytes(), "export differs"
M.close(), M0.close()
print("parity: ok (13 column ranges, export bit for bit, SpMV to 1e-13)")
ctx.close()
# ---- time ----------------------------------------------------------------------------------------
out = []
SWEEP = [(0, None, None, None), (1, "64", "plain", "0"), (1, "64", "plain", "32768"), (1, "64", "grouped", "32768"),
         (1, "64", "grouped", "1024"), (1, "48", "grouped", "32768"), (1, "32", "grouped", "32768"),
         (1, "100", "grouped", "32768")]
if os.environ.get("COLBLOCK_SWEEP"):
    SWEEP = [tuple(None if f == "-" else f for f in t.split(":")) for t in os.environ["COLBLOCK_SWEEP"].split(",")]
    SWEEP = [(int(t[0]),) + t[1:] for t in SWEEP]
for blocked, mb, kern, sigma in SWEEP:
    env = dict(os.environ)
    if mb:
        env["B200_COL_BLOCK_MB"] = mb
    if kern:
        env["B200_COL_BLOCK_KERNEL"] = kern
    if sigma:
        env["B200_COL_BLOCK_SIGMA"] = sigma
    r = subprocess.run([sys.executable, __file__, "--time", str(rows), str(blocked)], env=env,
                       capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else \
        json.dumps({"error": (r.stderr or "")[-300:], "col_block_mb": mb})
    print(line, flush=True)
    out.append(json.loads(line))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "colblock_check.json"), "w"), indent=1)
