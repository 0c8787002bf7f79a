#!/usr/bin/env python
"""Row-partitioned GMRES on N GPUs of one box, checked against a direct solve (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank builds its plan, solves collectively (halo exchange + all-reduce, per-rank block-Jacobi ILU(0)), the
pieces are gathered and compared (FP64 cases) with the CPU ORACLE -- the reference's host GMRES (src/gmres.cu:1904-2066)
given the block-Jacobi ILU(0) of exactly the rank x tile partition the ranks used: same iteration count,
|resid - resid_oracle| <= 1e-10, ||x - x_oracle|| <= 1e-8 ||x_oracle|| -- and with scipy's sparse LU; the iteration count
is printed next to the one a single rank needs with the same tiles (iteration drift of the coarser splitting)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))


def rank_tile_blocks(bg, rp, ci, v, world, part, method, tile, nx, ys):
    """block_of_row of the GLOBAL matrix as the ranks build it: rank q's diagonal block cut into its own tiles
    (bg_dist_create: halo columns dropped, make_blocks with the rank's dims / METIS on the rank's pattern)."""
    import scipy.sparse as sp

    n = len(rp) - 1
    blk = np.zeros(n, np.int32)
    base = 0
    for q in range(world):
        plan = bg.DistPlan(rp, ci, v, world, q, part)
        a = plan.arrays()
        nl = plan.info.n_local
        L = sp.csr_matrix((a["val"], a["col"], a["rowptr"]), shape=(nl, nl + plan.info.n_halo))[:, :nl].tocsr()
        L.sort_indices()
        rows = max(32, tile * tile)
        nb = max(1, (nl + rows - 1) // rows)
        if method == "strips":
            nbq, bq = bg.host_block_partition(bg.PART_GRID2D, L.indptr, L.indices, nb, [nx, ys[q + 1] - ys[q], 1, tile, tile, 1])
        else:
            nbq, bq = bg.host_block_partition(bg.PART_METIS, L.indptr, L.indices, nb)
        blk[a["l2g"]] = base + bq
        base += nbq
        plan.close()
    return blk


def main():
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    import torch.distributed as dist

    import b200gmres as bg
    from b200gmres import synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # (nx, ny, tile, partition, vector dtype): strips = contiguous rows (<= 2 neighbours per rank), metis = k-way
    # parts of the off-diagonal graph (several neighbours per rank once world > 2), f32 = the reference's REAL
    cases = ((96, 80, 16, "strips", "f64", "cgs"), (300, 280, 32, "strips", "f64", "cgs"), (96, 80, 16, "metis", "f64", "cgs"),
             (257, 199, 32, "metis", "f64", "cgs"), (96, 80, 16, "strips", "f32", "cgs"),
             # strict MGS: i+1 all-reduces per iteration through dots_kernel; CGS2: two all-reduces inside one orth_kernel
             (96, 80, 16, "strips", "f64", "mgs"), (96, 80, 16, "metis", "f64", "cgs2"))
    for (nx, ny, tile, method, dtn, orthn) in cases:
        orth = {"mgs": bg.ORTH_MGS, "cgs": bg.ORTH_CGS_ADAPTIVE, "cgs2": bg.ORTH_CGS2}[orthn]
        dtc, npdt, tdt = (bg.F32, np.float32, torch.float32) if dtn == "f32" else (bg.F64, np.float64, torch.float64)
        tol = 1e-5 if dtn == "f32" else 1e-8
        rp, ci, v, b = synth.pg2d(nx, ny, seed=7)
        n = len(rp) - 1
        ys = None
        if method == "strips":
            ys = [(r * ny) // world for r in range(world + 1)]
            part = np.repeat(np.arange(world, dtype=np.int32), np.diff(ys) * nx)
            o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=tile * tile, block_method=bg.PART_GRID2D,
                                dims=[nx, ys[rank + 1] - ys[rank], 1, tile, tile, 1])
        else:
            _, part = bg.host_block_partition(bg.PART_METIS, rp, ci, world)
            o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=tile * tile, block_method=bg.PART_METIS)
        plan = bg.DistPlan(rp, ci, v, world, rank, part)
        ids = [bg.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx = bg.Context(local)
        D = bg.Dist(ctx, plan, ids[0], dtype=dtc, m=32, opts=o)
        a = plan.arrays()
        db = torch.from_numpy(np.ascontiguousarray(b[a["l2g"]]).astype(npdt)).to(dev)
        dx = torch.zeros(plan.info.n_local, dtype=tdt, device=dev)
        torch.cuda.synchronize()
        opts = bg.gmres_opts(m=32, tol=tol, max_iter=3000, reorth_eta=0.1, orth=orth)
        # both transports: our peer-memory kernels (default) and the NCCL baseline; second solve of each is the warm time
        per_transport = {}
        avail = bool(D.info().p2p_available)
        for tname, tr in (("nccl", bg.TRANSPORT_NCCL), ("p2p", bg.TRANSPORT_P2P)):
            if tr == bg.TRANSPORT_P2P and not avail:
                continue
            D.set_transport(tr)
            for rep in range(2):
                dx.zero_()
                torch.cuda.synchronize()
                dist.barrier()
                r = D.solve(db.data_ptr(), dx.data_ptr(), opts)
                ctx.sync()
            per_transport[tname] = dict(iters=int(r.iters), solve_ms=round(r.solve_ms, 3), resid=float(r.resid))
        x = np.zeros(n)
        for q in range(world):
            sz = int((part == q).sum())
            t = dx.clone() if q == rank else torch.zeros(sz, dtype=tdt, device=dev)
            dist.broadcast(t, src=q)
            x[part == q] = t.cpu().numpy().astype(np.float64)
        info = D.info()
        if rank == 0:
            A = sp.csr_matrix((v, ci, rp), shape=(n, n))
            xs = spla.spsolve(A.tocsc(), b)
            err = float(np.linalg.norm(x - xs) / np.linalg.norm(xs))
            res = float(np.linalg.norm(b - A @ x) / np.linalg.norm(b))
            # single-rank run with the same tile size for the drift report
            c1 = bg.Context(local)
            o1 = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=tile * tile,
                                 block_method=bg.PART_GRID2D if method == "strips" else bg.PART_METIS,
                                 dims=[nx, ny, 1, tile, tile, 1])
            A1 = bg.Matrix(c1, rp, ci, v, dtype=dtc)
            M1 = bg.Precond.ilu0(c1, rp, ci, v, dtype=dtc, opts=o1)
            S1 = bg.Solver(c1, n, 32, dtype=dtc)
            d1b = torch.from_numpy(b.astype(npdt)).to(dev)
            d1x = torch.zeros(n, dtype=tdt, device=dev)
            torch.cuda.synchronize()
            r1 = S1.solve(A1, M1, d1b.data_ptr(), d1x.data_ptr(), opts)
            lim = (5e-3, 1e-4) if dtn == "f32" else (1e-6, 1e-7)
            good = r.flag == 0 and err < lim[0] and res < lim[1] and len({t["iters"] for t in per_transport.values()}) == 1
            vs_oracle = None
            if dtn == "f64":
                # the oracle with the same rank x tile block-Jacobi ILU(0): north_star bars
                from oracle import pyoracle as O

                O.build()
                blk = rank_tile_blocks(bg, rp, ci, v, world, part, method, tile, nx, ys if method == "strips" else None)
                f = bg.host_ilu0(rp, ci, v, blk)
                pc = O.Precond(O.PC_ILU0, n, "f64", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
                oflag, xo, oit, ores, _ = O.gmres(rp, ci, v, b, np.zeros(n), 32, 3000, tol, pc)
                xerr = float(np.linalg.norm(x - xo) / np.linalg.norm(xo))
                vs_oracle = dict(iters_oracle=int(oit), x_vs_oracle=xerr, resid_diff=abs(float(r.resid) - float(ores)))
                good = good and r.iters == oit and xerr <= 1e-8 and abs(r.resid - ores) <= 1e-10
            ok = ok and good
            print(json.dumps(dict(case=f"pg2d {nx}x{ny} {method} {dtn} {orthn}", ranks=world, flag=r.flag, iters=r.iters, iters_1gpu=r1.iters, resid=r.resid,
                                  true_rel_residual=res, x_err_vs_direct=err, vs_oracle=vs_oracle, halo_bytes=int(info.halo_bytes_per_exchange),
                                  neighbors=int(info.n_neighbors), transport=("p2p" if info.transport == bg.TRANSPORT_P2P else "nccl"), transports=per_transport,
                                  solve_ms=round(r.solve_ms, 3),
                                  solve_ms_1gpu=round(r1.solve_ms, 3), ok=bool(good))), flush=True)
            S1.close(); M1.close(); A1.close(); c1.close()
        D.close(); plan.close(); ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
