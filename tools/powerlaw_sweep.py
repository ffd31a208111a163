"""BASELINE.json config 5: irregular-row CSR (power-law row lengths) SpMV
kernel-selection sweep.  One rank per GPU (torchrun for N > 1), each rank owns
its row block; the halo of this matrix is near all-to-all (half the columns
are uniform), so the multi-rank SpMV is exchange-then-multiply.

    python tools/powerlaw_sweep.py [rows] [seeds...]
    torchrun --nproc-per-node 8 tools/powerlaw_sweep.py 50000000 1

Prints one JSON line per (seed, selection): bins, padding, ms per SpMV (CUDA
events on the launch stream, max over ranks) and algorithmic GB/s summed over
ranks (12 nnz + 4 (n+1) + 16 n, local rows; padding and halo not counted).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from lsbench_b200 import abi

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
seeds = [int(a) for a in sys.argv[2:]] or [1]
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
nccl_id = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = torch.zeros(abi.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(abi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    nccl_id = bytes(t.cpu().numpy().tobytes())
ctx = abi.Context(local, rank, world, nccl_id)

SEL = [("auto", 0), ("sell_only", abi.MAT_FORCE_SELL), ("vector_only", abi.MAT_FORCE_VECTOR),
       ("auto_nosort", abi.MAT_NO_SORT)]
# column blocking (B200_MAT_COL_BLOCK; width from B200_COL_BLOCK_MB, default 64 MB of x)
SEL.insert(1, ("auto_colblock", abi.MAT_COL_BLOCK))
if os.environ.get("POWERLAW_SWEEP_ONLY"):
    SEL = [s_ for s_ in SEL if s_[0] in os.environ["POWERLAW_SWEEP_ONLY"].split(",")]
for seed in seeds:
    for name, fl in (SEL if seed == seeds[0] else SEL[:1]):
        t0 = time.time()
        try:
            M = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, rows, seed=seed, flags=fl)
        except abi.B200Error as e:
            if rank == 0:
                print(json.dumps({"seed": seed, "selection": name, "error": str(e)[:200]}))
            continue
        i = M.info()
        n = i.n_local
        sp, _ = M.algorithmic_bytes()
        dx, dy = ctx.array(n), ctx.array(n)
        dx.upload(np.random.default_rng(rank).standard_normal(n))
        ms = min(M.spmv_time(dx, dy, reps=20) for _ in range(3))
        tot_bytes, ms_max = float(sp), float(ms)
        if world > 1:
            v = torch.tensor([tot_bytes, 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(v)
            tot_bytes = float(v[0].item())
            m = torch.tensor([ms_max], dtype=torch.float64, device=dev)
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            ms_max = float(m.item())
        if rank == 0:
            print(json.dumps({
                "workload": "powerlaw:%d" % rows, "seed": seed, "selection": name, "n_gpus": world,
                "rank0": {"n_local": n, "nnz": i.nnz, "nnz_padded": i.nnz_padded,
                          "padding_pct": 100.0 * (i.nnz_padded - i.nnz) / max(i.nnz, 1),
                          "sell_rows": i.sell_rows, "sell_sigma": i.sell_sigma,
                          "sell_max_width": i.sell_max_width, "vec_rows": i.vec_rows,
                          "vec_nnz": i.vec_nnz, "long_rows": i.long_rows, "long_nnz": i.long_nnz,
                          "max_row_len": i.max_row_len, "n_halo": i.n_halo, "col_blocks": i.col_blocks,
                          "matrix_stream_bytes": i.matrix_stream_bytes,
                          "hist": list(i.hist)},
                "setup_s": time.time() - t0, "ms_per_spmv": ms_max,
                "algorithmic_bytes_all_ranks": tot_bytes,
                "algorithmic_gbs": tot_bytes / ms_max / 1e6}))
            sys.stdout.flush()
        dx.free(); dy.free()
        M.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
ctx.close()
