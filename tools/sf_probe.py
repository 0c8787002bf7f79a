#!/usr/bin/env python
"""Development aid: where the warps of trsv_syncfree spend their time, from %globaltimer stamps compiled in with
-DBG_SF_TIMING (make -C gpu-gmres_b200 dbg):   python tools/sf_probe.py [--nx 652] [--ilupp]
Per level: slices, first take, first / last publication; per slice: taken -> operands loaded -> dependencies arrived ->
published.  --ilupp: the split preconditioner on ILU++ factors (input from the reference's producer, oracle/_ref)."""
import argparse, os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import b200gmres as bg
bg.LIB_PATH = os.path.join(ROOT, "gpu-gmres_b200", "build", "libb200gmres_dbg.so")
import torch
from b200gmres import synth
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=652)
ap.add_argument("--ilupp", action="store_true")
ap.add_argument("--every", type=int, default=1, help="print every k-th level")
a = ap.parse_args()
CAP = 65536
rp, ci, v, b = synth.pg2d(a.nx, a.nx, seed=synth.SEED_BASE + 2)
n = len(rp) - 1
ctx = bg.Context(0)
L = bg.lib()
if a.ilupp:
    import scipy.sparse as sp
    from oracle import pyoracle as O
    Ac = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc(); Ac.sort_indices()
    fac = O.ilupp_factor(n, Ac.indptr, Ac.indices, Ac.data, threshold=1.2, mem_factor=1.0)
    keys = ("lp", "li", "lv", "up", "ui", "uv", "middle", "lscale", "rscale", "permRow", "permCol")
    M = bg.Precond.split(ctx, *[fac[k] for k in keys], dtype=bg.F64)
    f = {k: np.asarray(fac[k]) for k in keys[:6]}
    solves = (("L", bg.PC_LEFT, 1, "lp", "li", "lv"), ("U", bg.PC_RIGHT, 0, "up", "ui", "uv"))
else:
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=bg.precond_opts(lattice=0))
    f = bg.host_ilu0(rp, ci, v)
    solves = (("L+U (stamps of U)", bg.PC_FULL, 0, "up", "ui", "uv"),)
din = torch.from_numpy(b).cuda(); dout = torch.zeros(n, dtype=torch.float64, device="cuda")
for name, which, lower, kp, ki, kv in solves:
    lev = bg.host_trsv_levels(lower, f[kp], f[ki], f[kv])
    cnt = np.bincount(lev)
    spl = (cnt + 31) // 32
    first = np.concatenate([[0], np.cumsum(spl)])
    nsl = int(first[-1])
    for _ in range(3):
        M.apply(which, din.data_ptr(), dout.data_ptr())
    ctx.sync()
    L.bg_debug_sf_times(C.c_int(0), None)
    M.apply(which, din.data_ptr(), dout.data_ptr())
    ctx.sync()
    t = (C.c_ulonglong * (CAP * 5))()
    L.bg_debug_sf_times(C.c_int(-1), t)
    t = np.frombuffer(t, dtype=np.uint64).astype(np.int64).reshape(CAP, 5)[:min(nsl, CAP)]
    ok = t[:, 3] > 0
    base = t[ok][:, 0].min()
    print(f"== {name}: n={n} slices={nsl} levels={len(cnt)} stamped={int(ok.sum())} total {(t[ok][:,3].max()-base)/1e3:.1f} us, SMs {len(set(t[ok][:,4]))}")
    d01, d12, d23 = t[ok][:, 1] - t[ok][:, 0], t[ok][:, 2] - t[ok][:, 1], t[ok][:, 3] - t[ok][:, 2]
    for nm, d in (("taken->operands loaded", d01), ("operands->deps arrived", d12), ("deps->published", d23)):
        print(f"  {nm:26s} ns: median {np.median(d):9.0f}  p10 {np.percentile(d,10):9.0f}  p90 {np.percentile(d,90):9.0f}  max {d.max():9.0f}")
    prev_last = 0
    for l in range(len(cnt)):
        a0, a1 = int(first[l]), min(int(first[l + 1]), CAP)
        if a0 >= CAP: break
        r = t[a0:a1]
        last = int(r[:, 3].max() - base)
        if l % a.every == 0 or l == len(cnt) - 1:
            print(f"  level {l:4d}: {a1-a0:6d} slices  first take {int(r[:,0].min()-base):8d}  last take {int(r[:,0].max()-base):8d}  first pub {int(r[:,3].min()-base):8d}  last pub {last:8d}  (+{last-prev_last:6d})  median wait {int(np.median(r[:,2]-r[:,1])):7d}")
        prev_last = last
M.close(); ctx.close()
