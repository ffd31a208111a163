#!/usr/bin/env python
"""bench.py -- headline benchmark of the lsbench `--solver b200` hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload poisson27:512]

Metric (BASELINE.json): fp64 Jacobi-PCG time to ||b - A x|| / ||b|| <= 1e-10 on the
27-point 3D Poisson operator 512^3 (134 217 728 rows, 3 609 741 304 nnz), b[i] = i,
x0 = 0 (src/lsbench.c:157-160), at 1/2/4/8 B200, row-block partitioned.  One "step" =
one whole solve.  Strong scaling: the problem is fixed, ranks share it.  For N > 1
launch under torchrun, one rank per GPU.

Nothing inside the timed solves is instrumented: the per-kernel-class times of the
`roofline` block come from ONE extra, untimed solve (B200_PCG_TIME_KERNELS).

The JSON line also carries
  parity        iterations, recurrence and true residual, residual replacements, and a
                checksum of x (sum, 2-norm, x at 8 fixed rows) -- the same at every N
  roofline      the dominant kernel (SELL SpMV fused with p.Ap).  `frac` is the DRAM-side
                fraction: bytes the stored layout makes one launch move / launch time /
                measured copy peak; `frac_algorithmic` counts SURVEY 8d's algorithmic
                bytes instead (uniform slices store w deltas instead of 32 w columns,
                so that one can exceed 1)
  e2e           the same solve through b200_pcg_solve_host (the X_bench call shape):
                pinned host b and x0 in, x out, copies inside the timing
  N = 1 only    spmv_7pt_256, pcg_7pt_256 (BASELINE config 3), nek (config 2: on-chip and
                streaming b200 PCG, CPU direct stand-in, the reference's own cuSOLVER
                backend from oracle/_ref/driver_cusolver), uncompressed (explicit columns),
                values_f32 (the opt-in fp32-stored operator, lossless here, beside the headline)
  powerlaw_50m  BASELINE config 5, SpMV, at every N
  cpu_baseline  the CPU oracle's OpenMP Jacobi-PCG passes (oracle/, kind "port") over a
                real row slab of the same operator, scaled by rows and iterations
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOL = 1e-10
MAXIT = 20000
METRIC = "pcg_time_to_1e-10"
UNIT = "s"
# iterations the 27-point 512^3 solve takes (measured, identical on 1/2/4/8 GPUs and
# run to run); the reference arm cannot run the full solve and scales its sample to it
KNOWN_ITERS = {"poisson27:512": 1177, "poisson7:256": 1017}


def parse_workload(w):
    kind, size = w.split(":")
    return kind, int(size)


def nnz_of(kind, N):
    return (3 * N - 2) ** 3 if kind == "poisson27" else 7 * N ** 3 - 6 * N * N


def config_of(workload, iters):
    """what the line is quoted on: the same dict in both arms and at every N"""
    kind, N = parse_workload(workload)
    return {"workload": workload, "n": N ** 3, "nnz": nnz_of(kind, N), "tol": TOL,
            "rhs": "b[i]=i", "x0": "0", "iterations": iters,
            "l2": "inputs >> L2 (flush not needed: matrix 33 GB, vectors 1 GB each at N=1)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"], capture_output=True, text=True,
                    timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names)
                   if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(self.rows), "reasons": reasons}


# ---- the CPU legs: oracle port, OpenMP over every host core ---------------------------------
_SLAB = {}


def cpu_slab(kind, size, planes):
    """rows of `planes` z-planes out of the middle of the REAL operator (global column
    ids), generated once per process"""
    import orc
    key = (kind, size, planes)
    if key not in _SLAB:
        _SLAB.clear()
        gen = orc.gen_poisson27 if kind == "poisson27" else orc.gen_poisson7
        p0 = (size - planes) // 2
        row0 = size * size * p0
        t0 = time.perf_counter()
        _SLAB[key] = (gen(size, row0, row0 + size * size * planes), row0, time.perf_counter() - t0)
    return _SLAB[key]


def cpu_pcg_sample(kind, size, full_iters, planes, its):
    """One timing sample: `its` Jacobi-PCG iterations' worth of passes (SpMV + p.q, the
    x / r update with its two sums, the p update: oracle/krylov.c orc_pcg_slab_seconds,
    the loop body of orc_pcg_omp) over a slab of real rows of the real operator with
    every host core, scaled by rows and to the iteration count of the full solve."""
    import orc
    threads = orc.set_threads(0)   # every online core, whatever OMP_NUM_THREADS says
    planes = min(planes, size)
    M, row0, gen_s = cpu_slab(kind, size, planes)
    dt = orc.pcg_slab_seconds(M, size ** 3, row0, its)
    if dt <= 0:
        raise RuntimeError("oracle slab timing failed")
    n_full = size ** 3
    est = dt / its * (n_full / M.n) * full_iters
    return {"value": est, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "oracle OpenMP Jacobi-PCG passes, %d iterations over rows of %d z-planes of the real "
                      "%s %d^3 operator (%d rows, %d nnz, global columns) in %.3f s with %d threads; scaled "
                      "by rows (x%.1f) and to %d iterations" % (its, planes, kind, size, M.n, M.nnz, dt,
                                                                threads, n_full / M.n, full_iters),
            "sample_seconds": dt, "slab_generation_seconds": gen_s,
            "ns_per_row_iteration": dt / its / M.n * 1e9}


def run_reference(args, out):
    """--impl reference: the reference's CPU path for this metric.  Its own solve is
    CHOLMOD (src/cholmod-impl.h:58-63), which cannot be built offline and could not
    factor a 134 M-row grid; the arm times the oracle port of the same Jacobi-PCG with
    every host thread.  Protocol as src/cholmod-impl.h:45-63: set-up (here: generating
    the slab) untimed, `warmup` untimed samples, `steps` timed ones; each step is a
    bounded sample -- a few real iterations over 64 real z-planes (1/8 of the rows) --
    scaled to the full solve.  Under torchrun rank 0 alone runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, size = parse_workload(args.workload)
    full_iters = args.ref_iters or KNOWN_ITERS.get(args.workload, 1000)
    vals, info = [], None
    for i in range(args.warmup + args.steps):
        info = cpu_pcg_sample(kind, size, full_iters, args.cpu_planes, args.cpu_its)
        if i >= args.warmup:
            vals.append(info["value"])
    v = sum(vals) / len(vals)
    info["value"] = v
    info["spread"] = [min(vals), max(vals)]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload, full_iters),
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "value = seconds one full solve would take on this host's cores: measured seconds per "
                    "iteration per row on real rows of the operator x rows x iterations (the solve itself "
                    "would take minutes per step)"}
    out.emit(json.dumps(line))
    return 0


class OneLineStdout:
    """The contract is ONE JSON line on stdout.  Libraries loaded later (NCCL
    prints its version line there) write to file descriptor 1, so fd 1 is
    pointed at stderr for the duration of the run and the line goes to the
    saved descriptor."""

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.fd, (text + "\n").encode())


def main():
    out = OneLineStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="poisson27:512")
    ap.add_argument("--ref-iters", type=int, default=0,
                    help="iterations of the full solve the reference arm scales to (0 = the measured count)")
    ap.add_argument("--cpu-planes", type=int, default=64, help="z-planes of the CPU timing slab")
    ap.add_argument("--cpu-its", type=int, default=5, help="iterations per CPU timing sample")
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed end-to-end solves (at most --steps)")
    ap.add_argument("--no-extras", action="store_true",
                    help="only the headline solve (no config 2/3/5 legs, no uncompressed leg)")
    ap.add_argument("--no-compress", action="store_true",
                    help="keep one explicit u32 column per entry (B200_MAT_NO_COMPRESS)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the opt-in variant timed beside the headline: fp32-stored values "
                         "(B200_MAT_VALUES_F32, lossless on the stencils: same bits, fewer bytes)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out)

    import numpy as np
    import torch
    import torch.distributed as dist
    from lsbench_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        t = torch.zeros(abi.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(abi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        nccl_id = bytes(t.cpu().numpy().tobytes())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(vs):
        t = torch.tensor(list(vs), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    kind, size = parse_workload(args.workload)
    gen = {"poisson7": abi.GEN_POISSON7, "poisson27": abi.GEN_POISSON27}[kind]
    ctx = abi.Context(local, rank, world, nccl_id)
    # an explicit (non-default) stream: every torch op on the solve's vectors and every
    # kernel of the library are ordered in it, and the CUDA events below bracket both
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    assert stream.cuda_stream != 0
    t_setup = time.perf_counter()
    mflags = abi.MAT_NO_COMPRESS if args.no_compress else 0
    M = abi.Matrix.generate(ctx, gen, size, flags=mflags)
    info = M.info()
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    n = info.n_local
    spmv_bytes, iter_bytes = M.algorithmic_bytes()

    # inputs resident in HBM for `value`; pinned host copies for `e2e`
    b_host = torch.arange(info.row_begin, info.row_begin + n, dtype=torch.float64).pin_memory()
    x_host = torch.zeros(n, dtype=torch.float64).pin_memory()
    d_b = b_host.to(dev)
    d_x = torch.zeros(n, dtype=torch.float64, device=dev)
    flags = abi.PCG_NO_SMALL

    def solve_dev(fl=flags):
        d_x.zero_()  # x reset per trial (src/ginkgo.cpp:92)
        r, rc = M.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=fl)
        return r

    def solve_host():
        x_host.zero_()
        o, r = abi.PcgOpts(TOL, MAXIT, 0, flags), abi.PcgResult()
        rc = abi.load().b200_pcg_solve_host(M.h, b_host.data_ptr(), x_host.data_ptr(),
                                            abi.C.byref(o), abi.C.byref(r))
        if rc != 0:
            raise abi.B200Error(rc, abi.load().b200_last_error().decode())
        return r

    def timed(fn, steps):
        """EXACTLY `steps` calls between barrier+sync pairs; device time by CUDA
        events on the launch stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = [fn() for _ in range(steps)]
        e1.record(stream)
        barrier()
        return allmax(e0.elapsed_time(e1)), res

    for _ in range(args.warmup):
        solve_dev()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, results = timed(solve_dev, args.steps)
    clocks = sampler.summary()
    iters = results[-1].iters
    assert all(r.iters == iters and r.status == 0 for r in results), \
        [(r.iters, r.status) for r in results]
    launches = sum(r.kernel_launches for r in results)
    sec_per_solve = ms_total / 1e3 / args.steps

    # ---- parity: what the solve returned, in a form that can be compared across N ---------
    res = results[-1]
    fixed_rows = [(k * (info.n_global - 1)) // 7 for k in range(8)]
    mine = [float(d_x[g - info.row_begin].item()) if info.row_begin <= g < info.row_begin + n else 0.0
            for g in fixed_rows]
    sums = allsum([float(d_x.sum().item()), float((d_x * d_x).sum().item())] + mine)
    parity = {"iterations": iters, "status": res.status, "relres": res.relres,
              "true_relres": res.true_relres, "tol": TOL, "meets_bar": bool(res.true_relres <= TOL),
              "replacements": res.replacements,
              "x_sum": sums[0], "x_norm2": sums[1] ** 0.5,
              "x_at_rows": dict(zip([str(g) for g in fixed_rows], sums[2:])),
              "note": "true_relres = ||b - A x|| / ||b|| recomputed by the solver at exit (all ranks); "
                      "x_sum / x_norm2 / x_at_rows must agree across N to rounding"}

    # ---- per-kernel-class times: one extra solve, outside every timed region ----------------
    rk = solve_dev(flags | abi.PCG_TIME_KERNELS)
    torch.cuda.synchronize()
    spmv_ms, upd_ms, pupd_ms = rk.spmv_ms, rk.update_ms, rk.pupdate_ms

    # ---- end to end through the host-buffer entry point -----------------------------------
    solve_host()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ms_e2e, res_e2e = timed(solve_host, e2e_steps)
    e2e_val = ms_e2e / 1e3 / e2e_steps

    peak, peak_src = peaks()
    compressed = info.sell_uniform_slices > 0
    # bytes one launch must move with the layout actually stored (index-compressed
    # or not): matrix streams + x read once + y written once
    stored_bytes = info.matrix_stream_bytes + 16 * n
    kernel_key = "k_spmv_sellc_dot" if compressed else "k_spmv_sell_dot"
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            tj = json.load(open(tp)).get(args.workload, {})
            traffic = tj.get(kernel_key)
            traffic_src = ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum per launch "
                           "from the committed ncu capture " + tj.get("source_compressed" if compressed else "source", ""))
        except Exception:
            traffic = None
    roof = None
    if spmv_ms > 0:
        sg = stored_bytes / (spmv_ms * 1e-3) / 1e9
        ag = spmv_bytes / (spmv_ms * 1e-3) / 1e9
        roof = {"bound": "hbm",
                "kernel": "SELL SpMV fused with p.Ap (q = A p, p.q): k_spmv_sellc (index-compressed) / k_spmv_sell",
                "achieved": sg, "peak": peak, "unit": "GB/s", "frac": sg / peak,
                "bytes_per_launch": stored_bytes,
                "achieved_algorithmic": ag, "frac_algorithmic": ag / peak,
                "frac_algorithmic_of_nominal_8TBs": ag / 8000.0,
                "algorithmic_bytes_per_launch": spmv_bytes,
                "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "launch_ms": spmv_ms,
                "index_compressed_slices": "%d/%d" % (info.sell_uniform_slices, info.sell_slices),
                "timing": "CUDA events around the kernel class in the first 32 iterations of one extra, "
                          "untimed solve (rank 0's launches)",
                "note": "achieved/frac: bytes the STORED layout makes one launch move (matrix streams as "
                        "stored + x once + y once) -- the DRAM-side figure; *_algorithmic: SURVEY 8d's "
                        "12 nnz + 4 (n+1) + 16 n, which the index-compressed layout undercuts"}

    extra = {}
    if not args.no_extras:
        extra.update(extras_all_ranks(args, abi, torch, dist, ctx, dev, world, rank, stream, peak, allmax, allsum))
    if rank == 0 and world == 1 and not args.no_extras:
        M.close()
        M = None
        extra.update(extras_single_gpu(args, abi, torch, ctx, dev, stream, peak, gen, size, mflags, flags,
                                       d_b, d_x, spmv_bytes, iter_bytes, compressed))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_pcg_sample(kind, size, iters, min(args.cpu_planes, 32), args.cpu_its)
        except Exception as e:  # the checker is absent: say so, do not fail the GPU line
            cpu = {"error": str(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": sec_per_solve, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_solve * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload, iters),
            "run": {"parallelism": "row-block x%d" % world, "nnz_rank0": info.nnz,
                    "halo": os.environ.get("B200_HALO", "peer memory where it can be mapped"),
                    "allreduce": os.environ.get("B200_ALLREDUCE", "peer memory where it can be mapped"),
                    "device_GB_rank0": info.device_bytes / 1e9, "setup_s": t_setup,
                    "stream": "explicit torch.cuda.Stream shared with the library"},
            "parity": parity,
            "pcg": {"iterations": iters, "relres": res.relres, "true_relres": res.true_relres,
                    "replacements": res.replacements,
                    "ms_per_iteration": sec_per_solve * 1e3 / max(iters, 1),
                    "algorithmic_gbs_per_gpu": iter_bytes * iters / sec_per_solve / 1e9,
                    "kernel_ms": {"spmv_dot": spmv_ms, "update": upd_ms, "pupdate": pupd_ms},
                    "kernel_ms_source": "one extra untimed solve"},
            "roofline": roof,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8 * world,
                    "d2h_bytes_per_step": n * 8 * world, "steps": e2e_steps,
                    "true_relres": res_e2e[-1].true_relres, "iterations": res_e2e[-1].iters,
                    "call": "b200_pcg_solve_host (pinned host b, x0 in; x out)"},
            "gpu_launches": launches, "clocks": clocks,
        }
        line.update(extra)
        if cpu:
            line["cpu_baseline"] = cpu
        out.emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def extras_all_ranks(args, abi, torch, dist, ctx, dev, world, rank, stream, peak, allmax, allsum):
    """BASELINE.json config 5 at every N: SpMV of the 50 M-row power-law operator."""
    import numpy as np
    out = {}
    rows = 50_000_000
    legs = [("auto", 0), ("column_blocked", abi.MAT_COL_BLOCK)]
    res = {}
    for label, fl in legs:
        try:
            t0 = time.perf_counter()
            P = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, rows, seed=1, flags=fl)
            i = P.info()
            npl = i.n_local
            sp, _ = P.algorithmic_bytes()
            x = torch.randn(npl, dtype=torch.float64, device=dev)
            y = torch.empty(npl, dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            ms = allmax(min(P.spmv_time(x, y, reps=10) for _ in range(3)))
            tot = allsum([float(sp), float(i.nnz)])
            res[label] = {"ms_per_spmv": ms, "algorithmic_gbs": tot[0] / ms / 1e6,
                          "frac_algorithmic": tot[0] / ms / 1e6 / (peak * world),
                          "nnz": int(tot[1]), "col_blocks": i.col_blocks, "n_halo_rank0": i.n_halo,
                          "matrix_stream_bytes_rank0": i.matrix_stream_bytes,
                          "setup_s": time.perf_counter() - t0}
            P.close()
            del x, y
        except Exception as e:
            res[label] = {"error": str(e)[:200]}
    res["workload"] = "powerlaw:%d seed 1 (mean row 18, max 65 536, half the columns uniform)" % rows
    res["traffic_note"] = ("DRAM traffic per SpMV on one GPU from committed ncu captures, not this run: auto 49 GB "
                           "(profiles/r01_*), column-blocked 28 GB (profiles/r02_powerlaw_colblock_ncu.txt)")
    out["powerlaw_50m"] = res
    return out


def extras_single_gpu(args, abi, torch, ctx, dev, stream, peak, gen, size, mflags, flags, d_b, d_x,
                      spmv_bytes, iter_bytes, compressed):
    import numpy as np
    extra = {}
    # ---- config 3: 7-point 256^3, SpMV stand-alone and PCG to 1e-10 -------------------------
    legs = [("spmv_7pt_256", mflags)]
    if not args.no_compress:
        legs.append(("spmv_7pt_256_uncompressed", abi.MAT_NO_COMPRESS))
    for label, fl in legs:
        M7 = abi.Matrix.generate(ctx, abi.GEN_POISSON7, 256, flags=fl)
        i7 = M7.info()
        n7 = i7.n_local
        x7 = torch.randn(n7, dtype=torch.float64, device=dev)
        y7 = torch.empty(n7, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        sb7, ib7 = M7.algorithmic_bytes()
        ms7 = min(M7.spmv_time(x7, y7, reps=50) for _ in range(3))
        st7 = i7.matrix_stream_bytes + 16 * n7
        extra[label] = {"ms": ms7, "gbs": sb7 / ms7 / 1e6,
                        "frac_of_measured": sb7 / ms7 / 1e6 / peak,
                        "frac_of_nominal_8TBs": sb7 / ms7 / 1e6 / 8000.0,
                        "algorithmic_bytes": sb7, "stored_bytes": st7,
                        "stored_gbs": st7 / ms7 / 1e6, "stored_frac_of_measured": st7 / ms7 / 1e6 / peak,
                        "index_compressed_slices": "%d/%d" % (i7.sell_uniform_slices, i7.sell_slices)}
        if label == "spmv_7pt_256":
            b7 = torch.arange(n7, dtype=torch.float64, device=dev)
            x7.zero_()
            M7.pcg(b7, x7, tol=TOL, maxit=MAXIT, flags=flags)
            best = None
            for _ in range(3):
                x7.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                r7, _ = M7.pcg(b7, x7, tol=TOL, maxit=MAXIT, flags=flags)
                e1.record(stream)
                torch.cuda.synchronize()
                s7 = e0.elapsed_time(e1) / 1e3
                best = s7 if best is None or s7 < best else best
            extra["pcg_7pt_256"] = {"value": best, "unit": UNIT, "iterations": r7.iters, "status": r7.status,
                                    "true_relres": r7.true_relres, "replacements": r7.replacements,
                                    "ms_per_iteration": best * 1e3 / max(r7.iters, 1),
                                    "algorithmic_gbs": ib7 * r7.iters / best / 1e9,
                                    "frac_algorithmic": ib7 * r7.iters / best / 1e9 / peak,
                                    "workload": "poisson7:256 (BASELINE.json config 3), b[i]=i, x0=0, tol 1e-10"}
            del b7
        M7.close()
        del x7, y7
    # ---- the headline solve with one explicit column per entry ------------------------------
    if compressed:
        Mu = abi.Matrix.generate(ctx, gen, size, flags=abi.MAT_NO_COMPRESS)
        d_x.zero_()
        Mu.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=flags)
        torch.cuda.synchronize()
        d_x.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ru, _ = Mu.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=flags)
        e1.record(stream)
        torch.cuda.synchronize()
        su = e0.elapsed_time(e1) / 1e3
        d_x.zero_()
        rt, _ = Mu.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=flags | abi.PCG_TIME_KERNELS)
        extra["uncompressed"] = {
            "value": su, "unit": UNIT, "iterations": ru.iters, "true_relres": ru.true_relres,
            "ms_per_iteration": su * 1e3 / max(ru.iters, 1),
            "algorithmic_gbs_per_gpu": iter_bytes * ru.iters / su / 1e9,
            "kernel_ms": {"spmv_dot": rt.spmv_ms, "update": rt.update_ms, "pupdate": rt.pupdate_ms},
            "spmv_achieved_gbs": spmv_bytes / (rt.spmv_ms * 1e-3) / 1e9 if rt.spmv_ms > 0 else None,
            "spmv_frac_of_measured": spmv_bytes / (rt.spmv_ms * 1e-3) / 1e9 / peak if rt.spmv_ms > 0 else None,
            "note": "B200_MAT_NO_COMPRESS: 12 B/nnz streams, the layout the algorithmic-byte count describes"}
        Mu.close()
    if not args.no_variants:
        Mv = abi.Matrix.generate(ctx, gen, size, flags=mflags | abi.MAT_VALUES_F32)
        iv = Mv.info()
        d_x.zero_()
        Mv.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=flags)
        torch.cuda.synchronize()
        d_x.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rv, _ = Mv.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=flags)
        e1.record(stream)
        torch.cuda.synchronize()
        sv = e0.elapsed_time(e1) / 1e3
        d_x.zero_()
        rvt, _ = Mv.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=flags | abi.PCG_TIME_KERNELS)
        extra["values_f32"] = {"value": sv, "unit": UNIT, "iterations": rv.iters, "true_relres": rv.true_relres,
                               "ms_per_iteration": sv * 1e3 / max(rv.iters, 1), "values_f32": iv.values_f32,
                               "matrix_stream_bytes": iv.matrix_stream_bytes,
                               "kernel_ms": {"spmv_dot": rvt.spmv_ms, "update": rvt.update_ms, "pupdate": rvt.pupdate_ms},
                               "note": "opt-in (--precision FP32), NOT the headline: SELL values stored as fp32, which "
                                       "is exact for this operator (values_f32 = 1: the fp64 copy is dropped) -- fp64 "
                                       "arithmetic on the widened values, x bit-identical to the headline solve"}
        Mv.close()
    # ---- config 2: the Nek coarse-grid matrices --------------------------------------------
    try:
        extra["nek"] = nek_legs(abi, ctx)
    except Exception as e:
        extra["nek"] = {"error": str(e)[:300]}
    return extra


def nek_file(name):
    """tests/data/<name>.txt, unpacking the xz copy the repository ships once"""
    plain = os.path.join(ROOT, "tests", "data", name + ".txt")
    if os.path.exists(plain):
        return plain
    import lzma
    import tempfile
    d = os.path.join(tempfile.gettempdir(), "lsbench_b200_data")
    os.makedirs(d, exist_ok=True)
    out = os.path.join(d, name + ".txt")
    if not os.path.exists(out):
        with lzma.open(plain + ".xz") as f, open(out + ".tmp", "wb") as g:
            g.write(f.read())
        os.replace(out + ".tmp", out)
    return out


NEK = ["tj7a_A_12", "tj7a_A_15", "tj7a_A_18", "xn3b_A_10", "xn3b_A_12", "xn3b_A_15", "xn3b_A_18"]


def nek_legs(abi, ctx, trials=100, cu_trials=3):
    """BASELINE.json config 2 the way the reference runs it -- `driver --solver X --matrix
    F --trials=T` (bin/driver.c), warm-up + timed loop inside X_bench (src/cholmod-impl.h:
    45-63), `elapsed` from the CSV row: this repository's driver with --solver b200 (on-chip
    kernel with its default block-Jacobi preconditioner, LSBENCH_B200_PCG=jacobi for plain
    Jacobi on the same kernel, and LSBENCH_B200_PCG=stream for the streaming kernels), the reference's own
    cuSOLVER backend UNMODIFIED (oracle/_ref/driver_cusolver, src/cusparse.c:181-209; it
    refactors in every call), and the CPU direct-solve stand-in for the CHOLMOD backend
    (oracle LDL^T with RCM, factor untimed, one core).  x is checked against the direct
    solve (tests/golden/direct.npz) at the 1e-8 bar."""
    import numpy as np
    drv = os.path.join(ROOT, "lsbench_b200", "host", "_build", "driver")
    cu_drv = os.path.join(ROOT, "oracle", "_ref", "driver_cusolver")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "direct.npz"))
    if not os.path.exists(drv):
        return {"error": "lsbench_b200/host/_build/driver is not built"}

    def csv_rows(text):
        lines = text.splitlines()
        row = ext = None
        for j, ln in enumerate(lines):
            if ln.startswith("===matrix,") and j + 1 < len(lines):
                row = lines[j + 1].split(",")
            if ln.startswith("===b200:") and j + 1 < len(lines):
                ext = lines[j + 1].split(",")
        return row, ext

    rows = {}
    for name in NEK:
        path = nek_file(name)
        o = {}
        for label, env in (("onchip", {}), ("onchip_jacobi", {"LSBENCH_B200_PCG": "jacobi"}),
                           ("streaming", {"LSBENCH_B200_PCG": "stream"})):
            xf = os.path.join(os.path.dirname(path), name + ".%s.x" % label)
            e = dict(os.environ)
            e.update(env)
            r = subprocess.run([drv, "--solver", "b200", "--matrix", path, "--trials=%d" % trials,
                                "--dump-x", xf], capture_output=True, text=True, timeout=600, env=e)
            row, ext = csv_rows(r.stdout)
            if r.returncode != 0 or not row or not ext:
                o[label] = {"error": (r.stderr or r.stdout)[-200:]}
                continue
            x = np.fromfile(xf)
            o["n"], o["nnz"] = int(row[1]), int(row[2])
            o[label] = {"ms_per_solve": float(row[6]) / trials * 1e3, "trials": trials,
                        "iterations": int(ext[1]), "status": int(ext[2]), "true_relres": float(ext[4]),
                        "path": int(ext[7]), "block_jacobi": int(ext[8]) if len(ext) > 8 else 0,
                        "rel_diff_direct": float(np.linalg.norm(x - gold[name]) / np.linalg.norm(gold[name]))}
        o["ref_cusolver"] = None
        if os.path.exists(cu_drv):
            try:
                rr = subprocess.run([cu_drv, "--solver", "cusolver", "--matrix", path,
                                     "--trials=%d" % cu_trials], capture_output=True, text=True, timeout=300)
                row, _ = csv_rows(rr.stdout)
                if row:
                    o["ref_cusolver"] = {"ms_per_solve": float(row[6]) / cu_trials * 1e3, "trials": cu_trials,
                                         "kind": "the reference's src/cusparse.c (cusolverSpDcsrlsvchol, RCM), "
                                                 "unmodified, its own clock() figure"}
                else:
                    o["ref_cusolver"] = {"error": (rr.stderr or rr.stdout)[-160:]}
            except Exception as ex:
                o["ref_cusolver"] = {"error": str(ex)[:120]}
        try:   # the checker's direct solve as the CPU baseline beside it (kind "port")
            import orc
            A = orc.matrix_read(path)
            Mo = orc.op_upper_mirror(A)
            b = orc.rhs(Mo.n)
            F = orc.Ldlt(Mo, 1)
            F.solve(b)
            t0 = time.perf_counter()
            for _ in range(trials):
                F.solve(b)
            o["cpu_direct_standin"] = {"ms_per_solve": (time.perf_counter() - t0) / trials * 1e3, "cores": 1,
                                       "kind": "port: oracle LDL^T + RCM (CHOLMOD-equivalent restatement), "
                                               "factor untimed as src/cholmod-impl.h:25-26,45-63"}
        except Exception as ex:
            o["cpu_direct_standin"] = {"error": str(ex)[:120]}
        rows[name] = o
    return rows


if __name__ == "__main__":
    sys.exit(main())
