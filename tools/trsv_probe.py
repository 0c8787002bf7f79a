#!/usr/bin/env python
"""Development aid: where one warp of trsv_block_fused spends its time (descriptor fetch, right-hand side, chunk
loop, result store), from %globaltimer stamps compiled in with -DBG_TRSV_TIMING.

    make -C gpu-gmres_b200 dbg && python tools/trsv_probe.py 1118
"""
import os, sys, ctypes as C
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import b200gmres as bg
bg.LIB_PATH = os.path.join(ROOT, "gpu-gmres_b200", "build", "libb200gmres_dbg.so")
import torch
from b200gmres import synth
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1118
rp, ci, v, b = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 2)
n = len(rp) - 1
ctx = bg.Context(0)
A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=1024, block_method=bg.PART_GRID2D, dims=[nx, nx, 1, 32, 32, 1])
M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o)
S = bg.Solver(ctx, n, 32, dtype=bg.F64)
opts = bg.gmres_opts(m=32, tol=1e-8, max_iter=3000, reorth_eta=0.1)
db = torch.from_numpy(b).cuda(); dx = torch.zeros(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
L = bg.lib()
for probe in (0, 600, int(M.info().nblocks) - 1):
    L.bg_debug_trsv_times(C.c_int(probe), None)
    for _ in range(3):
        M.apply(bg.PC_FULL, db.data_ptr(), dx.data_ptr())   # one launch each (gather variant); the stamps of the last one are read
    ctx.sync()
    t = (C.c_ulonglong * 96)()
    L.bg_debug_trsv_times(C.c_int(-1), t)
    t = np.array(t[:], dtype=np.int64)
    t0 = t[0]
    nch = sum(1 for k in range(8, 96, 2) if t[k] > t0)
    print(f"probe block {probe}: desc+table {t[1]-t0} ns; issue {t[2]-t[1]}; rhs wait {t[3]-t[2]}; chunks {nch}")
    prev = t[3]
    row = []
    for c in range(nch):
        row.append((int(t[8+2*c]-prev), int(t[9+2*c]-t[8+2*c])))
        prev = t[9+2*c]
    print("   per chunk (wait ns, process ns):", row)
    print(f"   after loop {t[4]-t0} ns; store issued {t[5]-t0}; stored {t[6]-t0}; kernel entry -> block start {t0-t[7]} ns")

# the kernel alone, timed from the host (block-major vectors like inside the solver are not reachable through the
# public apply; this is the gather variant)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream = torch.cuda.ExternalStream(ctx.stream)
dy = torch.zeros_like(dx)
for _ in range(5): M.apply(bg.PC_FULL, db.data_ptr(), dy.data_ptr())
ctx.sync()
e0.record(stream)
for _ in range(50): M.apply(bg.PC_FULL, db.data_ptr(), dy.data_ptr())
e1.record(stream); torch.cuda.synchronize()
print("apply (gather variant), back to back:", e0.elapsed_time(e1) / 50 * 1e3, "us per launch")
