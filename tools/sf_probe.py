#!/usr/bin/env python
"""Development aid: where the warps of trsv_syncfree spend their time, from %globaltimer stamps compiled in with
-DBG_SF_TIMING (make -C gpu-gmres_b200 dbg):   python tools/sf_probe.py [nx] [q0]
Per slice: taken -> operands loaded -> dependencies arrived -> published; and the pace of the dependency front."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import b200gmres as bg
bg.LIB_PATH = os.path.join(ROOT, "gpu-gmres_b200", "build", "libb200gmres_dbg.so")
import torch
from b200gmres import synth
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1304
rp, ci, v, b = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 2)
n = len(rp) - 1
ctx = bg.Context(0)
M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=bg.precond_opts(lattice=0))
din = torch.from_numpy(b).cuda(); dout = torch.zeros(n, dtype=torch.float64, device="cuda")
L = bg.lib()
f = bg.host_ilu0(rp, ci, v)
lev = bg.host_trsv_levels(1, f["lp"], f["li"], f["lv"])
cnt = np.bincount(lev)
slices_per_level = (cnt + 31) // 32
first_slice = np.concatenate([[0], np.cumsum(slices_per_level)])
q0 = int(sys.argv[2]) if len(sys.argv) > 2 else int(first_slice[len(cnt) // 2])
for rep in range(3):
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
ctx.sync()
L.bg_debug_sf_times(C.c_int(q0), None)
M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())   # note: the U solve overwrites the stamps of the L solve for the same q
ctx.sync()
t = (C.c_ulonglong * (2048 * 5))()
L.bg_debug_sf_times(C.c_int(-1), t)
t = np.array(t[:], dtype=np.int64).reshape(2048, 5)
lm = len(cnt) // 2
b0 = t[t[:, 0] > 0][:, 0].min()
for l in range(lm, lm + 3):
    a0, a1 = int(first_slice[l]) - q0, int(first_slice[l + 1]) - q0
    rows = t[a0:a1]
    print(f"  level {l}: {a1-a0} slices; taken(min..max) {rows[:,0].min()-b0}..{rows[:,0].max()-b0}  deps arrived {sorted((rows[:,2]-b0).tolist())[:6]} .. max {rows[:,2].max()-b0}  published min {rows[:,3].min()-b0} max {rows[:,3].max()-b0}")
    print("     per slice (deps_arrived, published, sm):", [(int(r[2]-b0), int(r[3]-b0), int(r[4])) for r in rows[:12]])
ok = t[:, 3] > 0
t = t[ok]
base = t[:, 0].min()
print(f"nx={nx} q0={q0} slices stamped {len(t)}; SMs used {len(set(t[:,4]))}")
d01, d12, d23 = t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2]
for nm, d in (("taken->operands loaded", d01), ("operands->deps arrived", d12), ("deps->published", d23), ("taken->published", t[:, 3] - t[:, 0])):
    print(f"  {nm:26s} ns: median {np.median(d):9.0f}  p10 {np.percentile(d,10):9.0f}  p90 {np.percentile(d,90):9.0f}  max {d.max():9.0f}")
# pace of the front: publication time of consecutive slices
pub = np.sort(t[:, 3]) - base
print("  publication times of the first 48 stamped slices (ns from first take):", (pub[:48]).tolist())
# three consecutive levels in detail: when each slice saw its dependencies arrive (t2) and published (t3)
span = pub[-1] - pub[0]
print(f"  {len(pub)} slices published over {span} ns -> {span/len(pub):.0f} ns per slice")
