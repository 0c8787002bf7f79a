#!/usr/bin/env python
"""Time one preconditioner application (L and U solve) per schedule on the synthetic power grid / thermal grid:
    python tools/trsv_bench.py --nx 1304 [--reps 50]
reference operator (level: wavefront strips, sync-free slices) vs block-Jacobi tiles; checks the level modes against
each other.  Prints one JSON line per schedule."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import torch
import b200gmres as bg
from b200gmres import synth

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=1304)
ap.add_argument("--ny", type=int, default=0, help="grid height (default: square)")
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--th3d", type=int, default=0, help="edge of a 3-D thermal grid instead of the 2-D power grid")
ap.add_argument("--modes", default="lattice,syncfree,block")
ap.add_argument("--ilupp", action="store_true", help="split preconditioner on ILU++ factors of the grid (input from the reference's "
                "own producer, oracle/_ref/libilupp.so): times DevPrecond_left + DevPrecond_right = one L and one U solve")
a = ap.parse_args()
if a.ilupp:
    import scipy.sparse as sp
    from oracle import pyoracle as O
    rp, ci, v, b = synth.pg2d(a.nx, a.ny or a.nx, seed=synth.SEED_BASE + 2)
    n = len(rp) - 1
    Ac = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc(); Ac.sort_indices()
    t0 = time.perf_counter()
    fac = O.ilupp_factor(n, Ac.indptr, Ac.indices, Ac.data, threshold=1.2, mem_factor=1.0)
    t_f = time.perf_counter() - t0
    keys = ("lp", "li", "lv", "up", "ui", "uv", "middle", "lscale", "rscale", "permRow", "permCol")
    ctx = bg.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    t0 = time.perf_counter()
    M = bg.Precond.split(ctx, *[fac[k] for k in keys], dtype=bg.F64)
    setup = time.perf_counter() - t0
    info = M.info()
    din = torch.from_numpy(b).cuda(); dmid = torch.zeros(n, dtype=torch.float64, device="cuda"); dout = torch.zeros(n, dtype=torch.float64, device="cuda")
    def app():
        M.apply(bg.PC_LEFT, din.data_ptr(), dmid.data_ptr())
        M.apply(bg.PC_RIGHT, dmid.data_ptr(), dout.data_ptr())
    for _ in range(3):
        app()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.reps):
        app()
    e1.record(stream)
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / a.reps
    print(json.dumps({"mode": "ilupp-split", "n": n, "nx": a.nx, "sf_depth": os.environ.get("BG_SF_DEPTH"), "levels": [info.levelsL, info.levelsU],
                      "nnzL": len(fac["lv"]), "nnzU": len(fac["uv"]), "us_per_L_plus_U": us, "ilupp_s": t_f, "setup_s": setup,
                      "stream_bytes": info.stream_bytes, "moved_GBps": (info.stream_bytes + 4 * 8 * n) / us / 1e3,
                      "checksum": float(np.abs(dout.cpu().numpy()).sum())}), flush=True)
    M.close(); ctx.close()
    sys.exit(0)
if a.th3d:
    rp, ci, v, b = synth.th3d(a.th3d, a.th3d, a.th3d, seed=synth.SEED_BASE + 3)
    dims, meth, rows = [a.th3d] * 3 + [8, 8, 8], bg.PART_GRID3D, 512
else:
    ny = a.ny or a.nx
    rp, ci, v, b = synth.pg2d(a.nx, ny, seed=synth.SEED_BASE + 2)
    dims, meth, rows = [a.nx, ny, 1, 32, 32, 1], bg.PART_GRID2D, 1024
n = len(rp) - 1
ctx = bg.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
din = torch.from_numpy(b).cuda(); dout = torch.zeros(n, dtype=torch.float64, device="cuda")
ref = None
for mode in a.modes.split(","):
    t0 = time.perf_counter()
    if mode == "block":
        o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=meth, block_rows=rows, dims=dims)
    else:
        o = bg.precond_opts(lattice=1 if mode == "lattice" else 0)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o)
    setup = time.perf_counter() - t0
    info = M.info()
    for _ in range(3):
        M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.reps):
        M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / a.reps
    x = dout.cpu().numpy()
    err = None
    if mode != "block":
        if ref is None:
            ref = x
        else:
            err = float(np.linalg.norm(x - ref) / np.linalg.norm(ref))
    moved = info.stream_bytes + 4 * 8 * n
    print(json.dumps({"mode": mode, "n": n, "nx": a.nx, "ny": a.ny or a.nx, "sf_depth": os.environ.get("BG_SF_DEPTH"), "lattice": info.lattice, "levelsL": info.levelsL, "us_per_apply": us, "setup_s": setup,
                      "stream_bytes": info.stream_bytes, "moved_GBps": moved / us / 1e3, "algorithmic_GBps": info.trsv_bytes / us / 1e3,
                      "vs_first_level_mode": err}), flush=True)
    M.close()
ctx.close()
