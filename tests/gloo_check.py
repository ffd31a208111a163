"""World-size-2 (or more) CPU worker for the N > 1 host logic, `gloo` backend:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port P tests/gloo_check.py

What runs on the GPU ranks as CUDA + NCCL (lsbench_b200/csrc/dist.cu) is
restated here rank for rank on the CPU with the oracle's row-block generators:
the row block each rank owns comes from the product's own b200_row_block (the
C ABI, pure host arithmetic), local/halo renumbering, interior/boundary split,
halo exchange by point-to-point messages, and Jacobi-PCG whose scalars are
all-reduced.  Checked against the serial oracle on the whole operator.
Prints "GLOO_CHECK OK" from rank 0.
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

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    def allreduce(vals):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t)
        return t.numpy().copy()

    for name, size, gen in (("poisson27", 20, orc.gen_poisson27), ("poisson7", 24, orc.gen_poisson7)):
        n = size ** 3
        # every rank sees the same cuts, they tile [0, n) and are 32-aligned inside
        cuts = [abi.row_block(n, k, world) for k in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[k][1] == cuts[k + 1][0] for k in range(world - 1))
        assert all(c[1] % 32 == 0 for c in cuts[:-1])
        r0, r1 = cuts[rank]
        nloc = r1 - r0
        Mo = gen(size, r0, r1)                       # my rows, global column ids
        cols = Mo.cols.astype(np.int64)
        remote = (cols < r0) | (cols >= r1)
        halo = np.unique(cols[remote])               # ascending global order = slot order
        owner = np.searchsorted(np.array([c[1] for c in cuts]), halo, side="right")
        local = np.where(remote, nloc + np.searchsorted(halo, cols), cols - r0)
        # interior = maximal middle run of rows without a halo column (dist.cu)
        row_has_halo = np.add.reduceat(remote.astype(np.int64), Mo.offs[:-1].astype(np.int64)) > 0
        clean = np.flatnonzero(~row_has_halo)
        assert clean.size and np.all(np.diff(clean) == 1), "interior must be one run for z-slabs"

        def exchange(x_loc):
            """x_ext = [owned | halo]: ask each owner for the entries I read"""
            want = [halo[owner == k] for k in range(world)]
            counts = torch.tensor([w.size for w in want], dtype=torch.int64)
            all_counts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(all_counts, counts)
            reqs, bufs = [], {}
            # send my wish lists, receive the others'
            asked = {}
            for k in range(world):
                if k == rank:
                    continue
                if want[k].size:
                    reqs.append(dist.isend(torch.from_numpy(want[k].copy()), k, tag=1))
                m = int(all_counts[k][rank])
                if m:
                    asked[k] = torch.zeros(m, dtype=torch.int64)
                    reqs.append(dist.irecv(asked[k], k, tag=1))
            for r in reqs:
                r.wait()
            reqs = []
            for k, idx in asked.items():             # pack + send (k_halo_pack)
                reqs.append(dist.isend(torch.from_numpy(x_loc[idx.numpy() - r0].copy()), k, tag=2))
            for k in range(world):
                if k != rank and want[k].size:
                    bufs[k] = torch.zeros(want[k].size, dtype=torch.float64)
                    reqs.append(dist.irecv(bufs[k], k, tag=2))
            for r in reqs:
                r.wait()
            x_ext = np.empty(nloc + halo.size)
            x_ext[:nloc] = x_loc
            for k, bsz in bufs.items():
                x_ext[nloc + np.flatnonzero(owner == k)] = bsz.numpy()
            return x_ext

        Ml = orc.Op(nloc, Mo.offs, local.astype(np.uint32), Mo.vals, ncols=nloc + halo.size)

        # ---- SpMV with halo exchange == the serial product, bit for bit (same row order)
        xg = np.random.default_rng(4).standard_normal(n)
        y = orc.spmv_fma(Ml, exchange(xg[r0:r1]))
        ys = [torch.zeros(c[1] - c[0], dtype=torch.float64) for c in cuts]
        dist.all_gather(ys, torch.from_numpy(y)) if len(set(c[1] - c[0] for c in cuts)) == 1 else None
        Mfull = gen(size)
        yfull = orc.spmv_fma(Mfull, xg)
        assert np.array_equal(y, yfull[r0:r1]), name

        # ---- Jacobi-PCG with all-reduced scalars (pcg.cu recurrences) ---------------------
        b = orc.rhs(n)[r0:r1]
        diag = np.array([Mo.vals[Mo.offs[i]:Mo.offs[i + 1]][cols[Mo.offs[i]:Mo.offs[i + 1]] == r0 + i][0]
                         for i in range(nloc)])
        dinv = 1.0 / diag
        x = np.zeros(nloc)
        r = b.copy()
        p = dinv * r
        rz, rr, bb = allreduce([r @ (dinv * r), r @ r, b @ b])
        its = 0
        while rr > 1e-20 * bb and its < 2000:
            q = orc.spmv_fma(Ml, exchange(p))
            pq = allreduce([p @ q])[0]
            alpha = rz / pq
            x += alpha * p
            r -= alpha * q
            rzn, rr = allreduce([r @ (dinv * r), r @ r])
            its += 1
            p = dinv * r + (rzn / rz) * p
            rz = rzn
        xs = [torch.zeros(c[1] - c[0], dtype=torch.float64) for c in cuts]
        sizes = [c[1] - c[0] for c in cuts]
        mx = max(sizes)
        pad = torch.zeros(mx, dtype=torch.float64)
        pad[:nloc] = torch.from_numpy(x)
        outs = [torch.zeros(mx, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(outs, pad)
        xfull = np.concatenate([o[:s].numpy() for o, s in zip(outs, sizes)])
        if rank == 0:
            xc, itc, _, _ = orc.pcg(Mfull, orc.rhs(n))
            assert abs(its - itc) <= 2, (its, itc)
            assert np.linalg.norm(xfull - xc) / np.linalg.norm(xc) <= 1e-8
            assert orc.true_relres(Mfull, orc.rhs(n), xfull) <= 1e-10
            print("gloo_check %s:%d ranks=%d rows=%s halo0=%d iters=%d (serial %d) ok"
                  % (name, size, world, sizes, halo.size, its, itc))
    dist.barrier()
    if rank == 0:
        print("GLOO_CHECK OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
