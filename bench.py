#!/usr/bin/env python
"""bench.py -- headline benchmark of the lsbench `--solver b200` hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload poisson27:512]

Metric (BASELINE.json): fp64 Jacobi-PCG time to ||r||/||b|| <= 1e-10 on the
27-point 3D Poisson operator 512^3 (134 217 728 rows, 3 609 741 304 nnz),
b[i] = i, x0 = 0 (src/lsbench.c:157-160), at 1/2/4/8 B200, row-block
partitioned.  One "step" = one whole solve.  Strong scaling: the problem is
fixed, ranks share it.  For N > 1 launch under torchrun, one rank per GPU.

The JSON line also carries
  roofline      the dominant kernel (SELL SpMV fused with p.Ap): algorithmic
                bytes per launch / mean launch duration, CUDA events on the
                launch stream, taken INSIDE the timed solves (first 32
                iterations of each) -- against MEASURED_PEAKS.json.  With the
                index-compressed layout (default) the kernel moves fewer bytes
                than the algorithmic count; stored_bytes / stored_frac give the
                DRAM-side view and `uncompressed` the same solve without it
  spmv_7pt_256  stand-alone fp64 SpMV GB/s on the 7-point 256^3 operator (the
                configuration the >= 75 % of 8 TB/s target is quoted on), N = 1
  e2e           the same solve through b200_pcg_solve_host (the X_bench call
                shape): pinned host b and x0 in, x out, copies inside the timing
  cpu_baseline  the CPU oracle's OpenMP Jacobi-PCG (oracle/, kind "port") on a
                bounded sample, scaled to the full solve
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


def parse_workload(w):
    kind, size = w.split(":")
    return kind, int(size)


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


_CPU_CACHE = {}


CPU_SAMPLE = {"n": None, "its": 600}   # overridden by --cpu-sample-n / --cpu-sample-its


def cpu_pcg_sample(kind, size, full_iters):
    """Oracle OpenMP Jacobi-PCG on a smaller cube of the same stencil; seconds
    per iteration per row, scaled to the full operator and iteration count."""
    import numpy as np
    import orc
    Ns = CPU_SAMPLE["n"] or (192 if kind == "poisson27" else 256)   # ~2.3 / ~1.5 GB of CSR: past the LLC
    gen = orc.gen_poisson27 if kind == "poisson27" else orc.gen_poisson7
    key = (kind, Ns)
    if key not in _CPU_CACHE:  # built once; every step re-times the iterations
        _CPU_CACHE.clear()
        _CPU_CACHE[key] = gen(Ns)
        orc.pcg(_CPU_CACHE[key], orc.rhs(_CPU_CACHE[key].n), maxit=2, omp=True)  # touch pages
    M = _CPU_CACHE[key]
    b = orc.rhs(M.n)
    its = CPU_SAMPLE["its"]   # default: ~10-30 s of CPU work on the GPU box's host cores
    t0 = time.perf_counter()
    _, it, _, _ = orc.pcg(M, b, tol=1e-30, maxit=its, omp=True)
    dt = time.perf_counter() - t0
    per_row_iter = dt / it / M.n
    n_full = size ** 3
    est = per_row_iter * n_full * full_iters
    return {"value": est, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": "oracle OpenMP Jacobi-PCG, %d iterations on %s %d^3 (%d rows) in %.2f s; "
                      "scaled by rows (x%.1f) and to %d iterations"
                      % (it, kind, Ns, M.n, dt, n_full / M.n, full_iters),
            "sample_seconds": dt}


def run_reference(args, out):
    """--impl reference: the reference's CPU path for this metric.  Its own
    solve is CHOLMOD (src/cholmod-impl.h:58-63), which cannot be built offline
    and cannot factor a 134 M-row grid anyway; the arm therefore times the
    oracle port of the same Jacobi-PCG with every host thread, each step a
    bounded sample scaled to the full solve."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, size = parse_workload(args.workload)
    full_iters = args.ref_iters
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_pcg_sample(kind, size, full_iters)
        if i >= args.warmup:
            vals.append(info["value"])
    v = sum(vals) / len(vals)
    info["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "tol": TOL, "rhs": "b[i]=i", "x0": "0",
                       "iterations_assumed": full_iters},
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    ap.add_argument("--ref-iters", type=int, default=1176,
                    help="iterations the full solve needs (measured on the GPU path)")
    ap.add_argument("--cpu-sample-n", type=int, default=0, help="grid edge of the CPU sample (0 = default)")
    ap.add_argument("--cpu-sample-its", type=int, default=600)
    ap.add_argument("--no-spmv", action="store_true")
    ap.add_argument("--no-compress", action="store_true",
                    help="keep one explicit u32 column per entry (B200_MAT_NO_COMPRESS)")
    ap.add_argument("--no-uncompressed-leg", action="store_true",
                    help="skip the extra NO_COMPRESS solve reported beside the headline")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--variants", action="store_true",
                    help="also time the opt-in variants beside the headline: fp32-stored values "
                         "(B200_MAT_VALUES_F32, lossless on the stencils) and single-reduction CG")
    args = ap.parse_args()
    CPU_SAMPLE["n"], CPU_SAMPLE["its"] = args.cpu_sample_n or None, args.cpu_sample_its
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

    kind, size = parse_workload(args.workload)
    gen = {"poisson7": abi.GEN_POISSON7, "poisson27": abi.GEN_POISSON27}[kind]
    ctx = abi.Context(local, rank, world, nccl_id)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
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
    flags = abi.PCG_NO_SMALL | abi.PCG_TIME_KERNELS

    def solve_dev():
        d_x.zero_()  # x reset per trial (src/ginkgo.cpp:92)
        r, rc = M.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=flags)
        return r

    def solve_host():
        x_host.zero_()
        o, r = abi.PcgOpts(TOL, MAXIT, 0, abi.PCG_NO_SMALL), abi.PcgResult()
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
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

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
    spmv_ms = sum(r.spmv_ms for r in results) / len(results)
    upd_ms = sum(r.update_ms for r in results) / len(results)
    pupd_ms = sum(r.pupdate_ms for r in results) / len(results)
    sec_per_solve = ms_total / 1e3 / args.steps

    # end to end through the host-buffer entry point
    solve_host()
    ms_e2e, res_e2e = timed(solve_host, args.steps)
    e2e_val = ms_e2e / 1e3 / args.steps

    peak, peak_src = peaks()
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
    compressed = info.sell_uniform_slices > 0
    kernel_key = "k_spmv_sellc_dot" if compressed else "k_spmv_sell_dot"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(kernel_key)
        except Exception:
            traffic = None
    # bytes one launch must move with the layout actually stored (index-compressed
    # or not): matrix streams + x read once + y written once
    stored_bytes = info.matrix_stream_bytes + 16 * n

    extra = {}
    if rank == 0 and world == 1 and not args.no_spmv:
        # the SpMV target configuration, stand-alone (matrix 1.5 GB >> 126 MB L2)
        M.close()
        legs = [("spmv_7pt_256", mflags)]
        if not args.no_compress:
            legs.append(("spmv_7pt_256_uncompressed", abi.MAT_NO_COMPRESS))
        for label, fl in legs:
            M7 = abi.Matrix.generate(ctx, abi.GEN_POISSON7, 256, flags=fl)
            i7 = M7.info()
            n7 = i7.n_local
            x7 = torch.randn(n7, dtype=torch.float64, device=dev)
            y7 = torch.empty(n7, dtype=torch.float64, device=dev)
            sb7, _ = M7.algorithmic_bytes()
            ms7 = min(M7.spmv_time(x7, y7, reps=50) for _ in range(3))
            extra[label] = {"ms": ms7, "gbs": sb7 / ms7 / 1e6,
                            "frac_of_measured": sb7 / ms7 / 1e6 / peak,
                            "frac_of_nominal_8TBs": sb7 / ms7 / 1e6 / 8000.0,
                            "algorithmic_bytes": sb7,
                            "stored_bytes": i7.matrix_stream_bytes + 16 * n7,
                            "index_compressed_slices": "%d/%d" % (i7.sell_uniform_slices, i7.sell_slices)}
            M7.close()
            del x7, y7
        if compressed and not args.no_uncompressed_leg:
            # the same solve with one explicit column per entry, beside the headline
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
            extra["uncompressed"] = {
                "value": su, "unit": UNIT, "iterations": ru.iters,
                "ms_per_iteration": su * 1e3 / max(ru.iters, 1),
                "algorithmic_gbs_per_gpu": iter_bytes * ru.iters / su / 1e9,
                "kernel_ms": {"spmv_dot": ru.spmv_ms, "update": ru.update_ms, "pupdate": ru.pupdate_ms},
                "spmv_achieved_gbs": spmv_bytes / (ru.spmv_ms * 1e-3) / 1e9 if ru.spmv_ms > 0 else None,
                "spmv_frac_of_measured": spmv_bytes / (ru.spmv_ms * 1e-3) / 1e9 / peak if ru.spmv_ms > 0 else None,
                "note": "B200_MAT_NO_COMPRESS: 12 B/nnz streams, the layout the algorithmic-byte count describes"}
            Mu.close()
        if args.variants:
            # opt-in variants, reported beside the headline, never as it
            def leg(mat_flags, pcg_flags, note):
                Mv = abi.Matrix.generate(ctx, gen, size, flags=mat_flags)
                iv = Mv.info()
                d_x.zero_()
                Mv.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=pcg_flags)
                torch.cuda.synchronize()
                d_x.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                rv, _ = Mv.pcg(d_b, d_x, tol=TOL, maxit=MAXIT, flags=pcg_flags)
                e1.record(stream)
                torch.cuda.synchronize()
                sv = e0.elapsed_time(e1) / 1e3
                out_v = {"value": sv, "unit": UNIT, "iterations": rv.iters, "status": rv.status,
                         "true_relres": rv.true_relres, "outer_iters": rv.outer_iters,
                         "ms_per_iteration": sv * 1e3 / max(rv.iters, 1),
                         "kernel_ms": {"spmv_dot": rv.spmv_ms, "update": rv.update_ms, "pupdate": rv.pupdate_ms},
                         "values_f32": iv.values_f32, "matrix_stream_bytes": iv.matrix_stream_bytes,
                         "note": note}
                Mv.close()
                return out_v
            extra["variants"] = {
                "values_f32": leg(mflags | abi.MAT_VALUES_F32, flags,
                                  "SELL values stored as fp32 (exact for this operator): fp64 arithmetic, same bits"),
                "single_reduction": leg(mflags, flags | abi.PCG_SINGLE_REDUCTION,
                                        "Chronopoulos-Gear CG: two kernels and one reduction point per iteration"),
                "values_f32+single_reduction": leg(mflags | abi.MAT_VALUES_F32,
                                                   flags | abi.PCG_SINGLE_REDUCTION, "both")}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_pcg_sample(kind, size, iters)

    if rank == 0:
        line = {
            "metric": METRIC, "value": sec_per_solve, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_solve * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": info.n_global, "nnz_local": info.nnz,
                       "tol": TOL, "rhs": "b[i]=i", "x0": "0", "iterations": iters,
                       "parallelism": "row-block x%d" % world,
                       "l2": "inputs >> L2 (matrix %.1f GB per GPU)" % (info.device_bytes / 1e9),
                       "setup_s": t_setup},
            "pcg": {"iterations": iters, "relres": results[-1].relres,
                    "true_relres": results[-1].true_relres,
                    "ms_per_iteration": sec_per_solve * 1e3 / max(iters, 1),
                    "algorithmic_gbs_per_gpu": iter_bytes * iters / sec_per_solve / 1e9,
                    "kernel_ms": {"spmv_dot": spmv_ms, "update": upd_ms, "pupdate": pupd_ms}},
            "roofline": {"bound": "hbm",
                         "kernel": ("k_spmv_sellc<dot>" if compressed else "k_spmv_sell<dot>")
                                   + " (q = A p, p.q)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None,
                         "frac_of_nominal_8TBs": achieved / 8000.0 if achieved else None,
                         "peak_source": peak_src, "traffic": traffic,
                         "algorithmic_bytes_per_launch": spmv_bytes,
                         "stored_bytes_per_launch": stored_bytes,
                         "stored_gbs": stored_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None,
                         "stored_frac": stored_bytes / (spmv_ms * 1e-3) / 1e9 / peak if spmv_ms > 0 else None,
                         "index_compressed_slices": "%d/%d" % (info.sell_uniform_slices, info.sell_slices),
                         "note": ("achieved counts the ALGORITHMIC bytes (12 nnz + 4 (n+1) + 16 n); "
                                  "uniform SELL slices store w column deltas instead of 32 w columns, "
                                  "so the kernel moves stored_bytes (< algorithmic) and frac can "
                                  "exceed the copy peak; stored_frac is the DRAM-side fraction"
                                  if compressed else "explicit u32 column per entry"),
                         "launch_ms": spmv_ms},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8 * world,
                    "d2h_bytes_per_step": n * 8 * world,
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


if __name__ == "__main__":
    sys.exit(main())
