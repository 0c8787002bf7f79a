#!/usr/bin/env python
"""Solve A x = b from files in the reference's formats on one B200 -- the flow of the thermal driver
`gmres_SpMV A.mtx [rhs] [out]` (src_thermal/main1.cu:126-405): read the triplet matrix, build the ILU(0) (or
DIAG / NONE) preconditioner, start from x = 1 like the driver does (:361), run GMRES(32), print flag, iterations and
the achieved tolerance in the driver's wording and write the solution one value per line.

    python tools/gmres_file.py A.mtx b.txt x.txt --precond ilu0 --dtype f32      # the reference's precision
    python tools/gmres_file.py A.bin --format csr_f64 --dtype f64 --tol 1e-8 --schedule block
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("matrix")
    ap.add_argument("rhs", nargs="?", help="one value per line; default b = A * ones like main1.cu when no rhs is given")
    ap.add_argument("out", nargs="?", default="xGPU.txt")
    ap.add_argument("--format", default="triplets", choices=["triplets", "csr_f32", "csr_f64"])
    ap.add_argument("--precond", default="ilu0", choices=["none", "diag", "ilu0"])
    ap.add_argument("--schedule", default="level", choices=["level", "block"], help="level = the reference's operator; block = block-Jacobi ILU(0)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--restart", type=int, default=32)          # src/defs.h:11
    ap.add_argument("--max-iter", type=int, default=60000)      # src/defs.h:11
    ap.add_argument("--tol", type=float, default=1e-6)          # src/defs.h:12
    args = ap.parse_args()

    import torch

    import b200gmres as bg

    fmt = {"triplets": bg.FILE_TRIPLETS, "csr_f32": bg.FILE_CSR_F32, "csr_f64": bg.FILE_CSR_F64}[args.format]
    rp, ci, v, ncols, empty = bg.host_read_matrix(args.matrix, fmt)
    n = len(rp) - 1
    if ncols != n:
        raise SystemExit(f"matrix is {n} x {ncols}: the solver needs a square system")
    if empty:
        raise SystemExit(f"{empty} rows have no entry (the reference's reader assumes none, SpMV_gen.cpp:157)")
    dtc, npdt, tdt = (bg.F32, np.float32, torch.float32) if args.dtype == "f32" else (bg.F64, np.float64, torch.float64)
    ctx = bg.Context(0)
    A = bg.Matrix(ctx, rp, ci, v, dtype=dtc)
    b = bg.host_read_vector(args.rhs, n, dtc) if args.rhs else None
    t0 = time.perf_counter()
    if args.precond == "none":
        M = bg.Precond.none(ctx, n, dtype=dtc)
    elif args.precond == "diag":
        M = bg.Precond.diag(ctx, rp, ci, v, dtype=dtc)
    else:
        o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_METIS, block_rows=1024) if args.schedule == "block" else None
        M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=dtc, opts=o)
    print(f"****** Time for preconditioning is {1e3 * (time.perf_counter() - t0):f}ms")
    S = bg.Solver(ctx, n, args.restart, dtype=dtc)
    dx = torch.ones(n, dtype=tdt, device="cuda")  # "guess soln" x = 1 (main1.cu:361)
    if b is None:
        db = torch.zeros(n, dtype=tdt, device="cuda")
        torch.cuda.synchronize()
        A.spmv(dx.data_ptr(), db.data_ptr())
        ctx.sync()
        dx.zero_()
    else:
        db = torch.from_numpy(b.astype(npdt)).cuda()
    torch.cuda.synchronize()
    orth = bg.ORTH_MGS if args.schedule == "level" else bg.ORTH_CGS_ADAPTIVE
    r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), bg.gmres_opts(m=args.restart, tol=args.tol, max_iter=args.max_iter, orth=orth))
    ctx.sync()
    print(f"GPU GMRES flag = {r.flag}")
    print(f"  iterations performed: {r.iters}")
    print(f"  tolerance achieved: {r.resid:8.6e}")
    print(f"  device time: {r.solve_ms:.3f} ms, {r.launches} kernel launches")
    bg.host_write_vector(args.out, dx.cpu().numpy())
    S.close(); M.close(); A.close(); ctx.close()
    return r.flag


if __name__ == "__main__":
    sys.exit(main())
