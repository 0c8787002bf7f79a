#!/usr/bin/env python
"""Development aid: where the four warps of one strip of trsv_lattice_global spend their cycles (clock64 accounting
compiled in with -DBG_SF_TIMING; make -C gpu-gmres_b200 dbg):   python tools/lg_probe.py [nx] [ny] [strip]"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import b200gmres as bg
bg.LIB_PATH = os.path.join(ROOT, "gpu-gmres_b200", "build", "libb200gmres_dbg.so")
import torch
from b200gmres import synth
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1304
ny = int(sys.argv[2]) if len(sys.argv) > 2 else nx
strip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rp, ci, v, b = synth.pg2d(nx, ny, seed=synth.SEED_BASE + 2)
n = len(rp) - 1
ctx = bg.Context(0)
M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64)
din = torch.from_numpy(b).cuda(); dout = torch.zeros(n, dtype=torch.float64, device="cuda")
L = bg.lib()
for _ in range(3):
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
ctx.sync()
L.bg_debug_lg_times(C.c_int(strip), None)
reps = 5
for _ in range(reps):
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
ctx.sync()
t = (C.c_longlong * 16)()
L.bg_debug_lg_times(C.c_int(-1), t)
t = np.array(t[:], dtype=np.float64) / (2 * reps)  # L and U launches both accumulate
steps = ny + 31
names = ["chain total", "chain waiting", " .. on loader", " .. on boundary", " .. on storer", "loader total", "loader flow control", "loader wait_group",
         "storer total", "storer waiting", "boundary total", "boundary polling"]
print(f"grid {nx}x{ny}, strip {strip}: {steps} steps per factor; cycles per step (mean of L and U):")
for k, nm in enumerate(names):
    print(f"  {nm:22s} {t[k]/steps:9.1f}")
