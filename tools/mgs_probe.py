#!/usr/bin/env python
"""Development aid: where one CTA of mgs_kernel spends a dot/axpy pair, from clock64 stamps compiled in with
-DBG_MGS_TIMING (make -C gpu-gmres_b200 dbg):   python tools/mgs_probe.py [nx] [iterations]"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import b200gmres as bg
bg.LIB_PATH = os.path.join(ROOT, "gpu-gmres_b200", "build", "libb200gmres_dbg.so")
import torch
from b200gmres import synth
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1304
its = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rp, ci, v, b = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 2)
n = len(rp) - 1
ctx = bg.Context(0)
A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID2D, block_rows=1024, dims=[nx, nx, 1, 32, 32, 1]))
S = bg.Solver(ctx, n, 32, dtype=bg.F64)
db = torch.from_numpy(b).cuda(); dx = torch.zeros(n, dtype=torch.float64, device="cuda")
L = bg.lib()
for probe in (0, 73):
    L.bg_debug_mgs_times(C.c_int(probe), None)
    dx.zero_(); torch.cuda.synchronize()
    r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), bg.gmres_opts(m=32, tol=1e-30, max_iter=its, orth=bg.ORTH_MGS))
    ctx.sync()
    t = (C.c_ulonglong * 200)()
    L.bg_debug_mgs_times(C.c_int(-1), t)
    t = np.frombuffer(t, dtype=np.uint64).reshape(40, 5)[:its].astype(np.int64)
    k = its  # the last launch ran inner iteration its-1: pairs 0..its-1, then the norm as "pair" its
    t = t[:k]
    print(f"n={n} CTA {probe}: last launch, {k} pairs; per pair (SM cycles, ~1.9 GHz): tile wait->local sum, ->stored, ->all read, ->pair done, pair total")
    prev = t[0, 0]
    for q in range(k):
        print(f"   pair {q:2d}: gap {t[q,0]-prev:6d}  local {t[q,1]-t[q,0]:6d}  store {t[q,2]-t[q,1]:6d}  poll {t[q,3]-t[q,2]:6d}  rest {t[q,4]-t[q,3]:6d}  total {t[q,4]-t[q,0]:6d}")
        prev = t[q, 4]
    print(f"   mean pair {np.mean(t[:,4]-t[:,0]):.0f} cycles, mean poll {np.mean(t[:,3]-t[:,2]):.0f} cycles")
S.close(); M.close(); A.close(); ctx.close()
