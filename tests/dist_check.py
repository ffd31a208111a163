"""Multi-rank parity worker, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P tests/dist_check.py

Checks, against the CPU oracle, for the row-block partitioned path
(lsbench_b200/csrc/dist.cu): partition + halo renumbering (exported local rows
== oracle rows with columns mapped back to global ids), SpMV with halo
exchange, PCG with all-reduced scalars (same x on every rank count, iteration
count reproducible).  Prints "DIST_CHECK OK" from rank 0.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import orc  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from lsbench_b200 import abi

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = torch.zeros(abi.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(abi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    ctx = abi.Context(local, rank, world, bytes(t.cpu().numpy().tobytes()))

    def gather(v):
        """all ranks' local slices -> the global vector, on every rank"""
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([v.size], dtype=torch.int64, device=dev))
        mx = max(int(s.item()) for s in sizes)
        buf = torch.zeros(mx, dtype=torch.float64, device=dev)
        buf[:v.size] = torch.from_numpy(v).to(dev)
        out = [torch.zeros(mx, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(out, buf)
        return np.concatenate([o[:int(s.item())].cpu().numpy() for o, s in zip(out, sizes)])

    cases = [("poisson27", 24, abi.GEN_POISSON27), ("poisson7", 40, abi.GEN_POISSON7),
             ("powerlaw", 40000, abi.GEN_POWERLAW)]
    for name, size, kind in cases:
        M = abi.Matrix.generate(ctx, kind, size, seed=9)
        i = M.info()
        r0, r1, n = i.row_begin, i.row_begin + i.n_local, i.n_global
        ref = {"poisson27": orc.gen_poisson27, "poisson7": orc.gen_poisson7}.get(name)
        Mo = ref(size, r0, r1) if ref else orc.gen_powerlaw(size, 9, r0, r1)
        # ---- partition / renumber: map local column ids back to global ----------------
        offs, cols, vals = M.export()
        halo = M.halo_cols()
        assert np.all(np.diff(halo.astype(np.int64)) > 0)
        assert np.all((halo < r0) | (halo >= r1))
        g = np.where(cols < i.n_local, cols.astype(np.int64) + r0,
                     halo[np.clip(cols.astype(np.int64) - i.n_local, 0, max(len(halo) - 1, 0))]
                     if len(halo) else 0)
        assert np.array_equal(offs, Mo.offs), name
        assert np.array_equal(g, Mo.cols.astype(np.int64)), name
        assert vals.tobytes() == Mo.vals.tobytes(), name
        assert sorted(set(Mo.cols[(Mo.cols < r0) | (Mo.cols >= r1)].tolist())) == halo.tolist()
        # ---- SpMV with halo exchange ---------------------------------------------------
        xg = np.random.default_rng(4).standard_normal(n)
        y = M.spmv_host(xg[r0:r1])
        yr, ya = orc.spmv(Mo, xg, want_abs=True)
        assert np.all(np.abs(y - yr) <= 1e-13 * ya + 1e-300), name
        if name != "powerlaw":
            assert np.array_equal(y, orc.spmv_fma(Mo, xg)), name
            # interior rows really have no halo column
            ib, ie = i.interior_begin, i.interior_end
            if ie > ib:
                assert np.all(cols[offs[ib]:offs[ie]] < i.n_local)
            # ---- PCG: same answer as the serial oracle on the whole grid ------------------
            b = orc.rhs(n)
            x, r, rc = M.pcg_host(b[r0:r1], tol=1e-10)
            assert rc == 0 and r.status == 0 and r.true_relres <= 1e-10
            assert r.relres <= 1e-10, r.relres
            x2, r2, _ = M.pcg_host(b[r0:r1], tol=1e-10)
            assert r2.iters == r.iters and x2.tobytes() == x.tobytes()
            xfull = gather(x)
            if rank == 0:
                Mfull = ref(size)
                xc, itc, _, _ = orc.pcg(Mfull, b)
                assert abs(r.iters - itc) <= 2, (r.iters, itc)
                assert np.linalg.norm(xfull - xc) / np.linalg.norm(xc) <= 1e-8
                assert orc.true_relres(Mfull, b, xfull) <= 1e-10
        if rank == 0:
            print("dist_check %s:%d ranks=%d halo=%d interior=[%d,%d) of %d ok"
                  % (name, size, world, i.n_halo, i.interior_begin, i.interior_end, i.n_local))
        M.close()

    # ---- column blocking on several ranks (BASELINE.json config 5): the ranges follow the global
    # column order [remote below | owned | remote above], the owned ranges are multiplied while
    # the halo travels; export and SpMV as for the plain layout --------------------------------
    os.environ["B200_COL_BLOCK_MB"] = "1"
    size = 1_000_000
    M = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, size, seed=9, flags=abi.MAT_COL_BLOCK)
    M0 = abi.Matrix.generate(ctx, abi.GEN_POWERLAW, size, seed=9)
    i = M.info()
    assert i.col_blocks >= 3, i.col_blocks
    r0, r1 = i.row_begin, i.row_begin + i.n_local
    o0, c0, v0 = M0.export()
    o1, c1, v1 = M.export()
    assert np.array_equal(o0, o1) and np.array_equal(c0, c1) and v0.tobytes() == v1.tobytes()
    assert np.array_equal(M.halo_cols(), M0.halo_cols())
    xg = np.random.default_rng(5).standard_normal(size)
    y, y0 = M.spmv_host(xg[r0:r1]), M0.spmv_host(xg[r0:r1])
    rows = min(i.n_local, 3000)
    Mo = orc.gen_powerlaw(size, 9, r0, r0 + rows)
    yr, ya = orc.spmv(Mo, xg, want_abs=True)
    assert np.all(np.abs(y[:rows] - yr) <= 1e-13 * ya + 1e-300)
    assert np.max(np.abs(y - y0)) <= 1e-11 * np.max(np.abs(y0))
    assert M.spmv_host(xg[r0:r1]).tobytes() == y.tobytes()
    if rank == 0:
        print("dist_check powerlaw:%d column-blocked ranks=%d ranges=%d halo=%d ok"
              % (size, world, i.col_blocks, i.n_halo))
    M.close(), M0.close()
    del os.environ["B200_COL_BLOCK_MB"]

    # a file matrix: every rank reads it, keeps its row block of the CHOLMOD operator
    A = orc.matrix_read(orc.matrix_path("tj7a_A_18"))
    M = abi.Matrix.from_csr(ctx, A.nrows, A.base, A.offs, A.cols, A.vals, abi.MAT_SYM_UPPER)
    i = M.info()
    Mo = orc.op_upper_mirror(A)
    b = orc.rhs(Mo.n)
    x, r, rc = M.pcg_host(b[i.row_begin:i.row_begin + i.n_local], tol=1e-10, maxit=5000)
    xfull = gather(x)
    if rank == 0:
        gold = np.load(os.path.join(HERE, "golden", "direct.npz"))["tj7a_A_18"]
        assert rc == 0 and np.linalg.norm(xfull - gold) / np.linalg.norm(gold) <= 1e-8
        assert orc.true_relres(Mo, b, xfull) <= 1e-10
        print("dist_check tj7a_A_18 ranks=%d iters=%d ok" % (world, r.iters))
    M.close()
    ctx.close()

    # ---- the ways the ranks can talk to each other must not change the answer -----------------
    # default: CG sums and the halo of p over peer memory, the chunk of iterations a CUDA graph;
    # B200_HALO=nccl: halo by ncclSend/Recv (no graph); B200_ALLREDUCE=nccl: sums by
    # ncclAllReduce as well.  A grid large
    # enough for several chunks of 32 iterations and a halo of two planes per neighbour.
    def fresh_ctx():
        t = torch.zeros(abi.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(abi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return abi.Context(local, rank, world, bytes(t.cpu().numpy().tobytes()))

    N = 64
    b = orc.rhs(N ** 3)
    got = {}
    for mode, env, pflags in (("peer", {}, 0), ("peer_nograph", {}, abi.PCG_NO_GRAPH),
                              ("halo_nccl", {"B200_HALO": "nccl"}, 0),
                              ("all_nccl", {"B200_ALLREDUCE": "nccl"}, 0)):
        for k in ("B200_HALO", "B200_ALLREDUCE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        c2 = fresh_ctx()
        M = abi.Matrix.generate(c2, abi.GEN_POISSON27, N)
        i = M.info()
        r0, r1 = i.row_begin, i.row_begin + i.n_local
        x, r, rc = M.pcg_host(b[r0:r1], tol=1e-10, flags=abi.PCG_NO_SMALL | pflags)
        x2, r2, _ = M.pcg_host(b[r0:r1], tol=1e-10, flags=abi.PCG_NO_SMALL | pflags)
        assert rc == 0 and r.status == 0 and r.true_relres <= 1e-10, (mode, rc, r.status, r.true_relres)
        assert r2.iters == r.iters and x2.tobytes() == x.tobytes(), mode     # reproducible
        got[mode] = (r.iters, gather(x), r.solve_ms)
        M.close()
        c2.close()
    for k in ("B200_HALO", "B200_ALLREDUCE"):
        os.environ.pop(k, None)
    it0, x0, _ = got["peer"]
    # the same sums in the same order: bit for bit, with or without the graph, whichever way the halo goes
    for mode in ("peer_nograph", "halo_nccl"):
        assert got[mode][0] == it0 and got[mode][1].tobytes() == x0.tobytes(), mode
    for mode in ("all_nccl",):
        assert abs(got[mode][0] - it0) <= 2, (mode, got[mode][0], it0)
        assert np.linalg.norm(got[mode][1] - x0) / np.linalg.norm(x0) <= 1e-9, mode
    if rank == 0:
        Mfull = orc.gen_poisson27(N)
        assert orc.true_relres(Mfull, b, x0) <= 1e-10
        print("dist_check modes ranks=%d: " % world
              + ", ".join("%s %d its %.2f ms" % (m, v[0], v[2]) for m, v in got.items()))
    dist.barrier()
    if rank == 0:
        print("DIST_CHECK OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
