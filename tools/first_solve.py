#!/usr/bin/env python
"""Wall-clock time to the FIRST solution of a 10 M-row power-grid system through the public API, call by call
(what a one-shot DC solve pays; bench.py's timed region starts after all of this): context, matrix upload + SpMV
storage, ILU(0) preconditioner, workspace, then three solves -- the first one also builds the block-major
reordering P A P^T the block-Jacobi path iterates on.

    python tools/first_solve.py [nx] [block|level]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import numpy as np, torch
import b200gmres as bg
from b200gmres import synth

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 3162
mode = sys.argv[2] if len(sys.argv) > 2 else "block"
rp, ci, v, b = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 2)
n = len(rp) - 1


def lap(label, fn):
    t = time.perf_counter()
    out = fn()
    print(f"{label:28s} {time.perf_counter() - t:7.3f} s", flush=True)
    return out


total = time.perf_counter()
ctx = lap("context", lambda: bg.Context(0))
A = lap("matrix (upload + SELL)", lambda: bg.Matrix(ctx, rp, ci, v, dtype=bg.F64))
if mode == "block":
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=1024, block_method=bg.PART_GRID2D, dims=[nx, nx, 1, 32, 32, 1])
else:
    o = bg.precond_opts(trsv_schedule=bg.TRSV_LEVEL)
M = lap(f"ILU(0) preconditioner ({mode})", lambda: bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o))
S = lap("workspace", lambda: bg.Solver(ctx, n, 32, dtype=bg.F64))
db = torch.from_numpy(b).cuda()
dx = torch.zeros(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
opts = bg.gmres_opts(m=32, tol=1e-8, max_iter=10000, orth=bg.ORTH_CGS_ADAPTIVE, reorth_eta=0.1)
for k in range(3):
    dx.zero_()
    torch.cuda.synchronize()

    def solve():
        r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), opts)
        ctx.sync()
        return r
    r = lap(f"solve {k}", solve)
    if k == 0:
        print(f"{'first solution after':28s} {time.perf_counter() - total:7.3f} s   ({r.iters} iterations, flag {r.flag}, residual {r.resid:.2e})")
