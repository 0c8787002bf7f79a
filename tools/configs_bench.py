#!/usr/bin/env python
"""The other single-GPU configurations of BASELINE.json on one B200 (bench.py carries the headline):

  thermal    configs[2]: 3D-IC thermal with liquid cooling, 7-point stencil, ~5M unknowns (171^3),
             GMRES(32)+ILU(0) (block-Jacobi tiles of 10x10x10 cells), DC-style solve of A = G + C/h
             (src_thermal/main1.cu:335-397 flow).
  transient  configs[3]: power-grid transient, 1M nodes (1000^2), `--steps` backward-Euler steps with the ILU
             factor built ONCE and every solve warm-started from the previous step
             (src/mna_solve_gpu_gmres.cpp:564-647: rhs = B u(t) + (C/h) x_{k-1}; x_h is never reset).
             The whole loop runs inside the library (bg_transient_run): PWL source evaluation, rhs = B u + (C/h) x
             by two device SpMVs, warm-started solve -- no host<->device vector traffic, one status per step.

    python tools/configs_bench.py --which thermal,transient --steps 1000

Prints one JSON line per configuration: time, iterations, achieved HBM GB/s against the algorithmic bytes
of SURVEY.md 8(d) and the true relative residual checked on the host.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def host_residual(rp, ci, v, x, b):
    import scipy.sparse as sp

    A = sp.csr_matrix((v, ci, rp), shape=(len(rp) - 1, len(rp) - 1))
    return float(np.linalg.norm(b - A @ x) / np.linalg.norm(b))


def thermal(args, bg, torch):
    from b200gmres import synth

    nx = args.thermal_n
    rp, ci, v, b = synth.th3d(nx, nx, nx, seed=synth.SEED_BASE + 3)
    n = len(rp) - 1
    ctx = bg.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
    t0 = time.perf_counter()
    t = args.thermal_tile
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=t * t * t, block_method=bg.PART_GRID3D, dims=[nx, nx, nx, t, t, t])
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o)
    setup = time.perf_counter() - t0
    S = bg.Solver(ctx, n, 32, dtype=bg.F64)
    opts = bg.gmres_opts(m=32, tol=args.tol, max_iter=10000, reorth_eta=0.1, orth={"cgs": bg.ORTH_CGS_ADAPTIVE, "cgs2": bg.ORTH_CGS2, "mgs": bg.ORTH_MGS}[args.orth])
    db = torch.from_numpy(b).cuda()
    dx = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ms = []
    for rep in range(3 + args.reps):
        with torch.cuda.stream(stream):
            dx.zero_()
        r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), opts)
        if rep >= 3:
            ms.append(r.solve_ms)
    with torch.cuda.stream(stream):
        dx.zero_()
    rprof, st = S.profile(A, M, db.data_ptr(), dx.data_ptr(), opts)
    ctx.sync()
    x = dx.cpu().numpy()
    tot = st.spmv_bytes + st.trsv_bytes + st.orth_bytes + st.other_bytes
    t_ms = float(np.mean(ms))
    info, pi = A.info(), M.info()
    line = {"config": f"BASELINE configs[2]: th3d {nx}^3 (7-point, micro-channel layers non-symmetric), A = G + C/h, GMRES(32)+ILU(0) block-Jacobi {t}^3 tiles",
            "n": n, "nnz": int(info.nnz), "solve_ms": t_ms, "iters": int(r.iters), "resid": float(r.resid), "true_rel_residual": host_residual(rp, ci, v, x, b),
            "us_per_iteration": 1e3 * t_ms / max(r.iters_done, 1), "achieved_GBps": tot / (t_ms * 1e-3) / 1e9, "frac_of_measured_hbm_peak": tot / (t_ms * 1e-3) / 1e9 / peak(),
            "spmv_format": int(info.format), "sell_padded_nnz": int(info.sell_padded_nnz), "blocks": int(pi.nblocks), "levels_per_block": [int(pi.levelsL), int(pi.levelsU)],
            "orth": args.orth, "profiled_ms": {"spmv": st.spmv_ms, "trsv": st.trsv_ms, "orth": st.orth_ms, "other": st.other_ms}, "setup_s": setup}
    line["orth"] = args.orth
    print(json.dumps(line), flush=True)
    S.close(); M.close(); A.close(); ctx.close()


def transient(args, bg, torch):
    from b200gmres import synth

    nx = args.transient_nx
    # A = G + C/h with C/h comparable to the conductances (h = 1e-11 s, C ~ 1-2 pF scaled)
    coh = 0.1
    rp, ci, v, b0 = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 4, c_over_h=coh)
    n = len(rp) - 1
    # recover the C/h diagonal the generator added (same seed stream): regenerate without it and subtract
    _, _, v0, _ = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 4, c_over_h=0.0)
    diag_pos = np.flatnonzero(ci == np.repeat(np.arange(n, dtype=np.int32), np.diff(rp)))
    c_h = v[diag_pos] - v0[diag_pos]
    ctx = bg.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
    t0 = time.perf_counter()
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=1024, block_method=bg.PART_GRID2D, dims=[nx, nx, 1, 32, 32, 1])
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o)  # factor built once, reused by every step
    setup = time.perf_counter() - t0
    S = bg.Solver(ctx, n, 32, dtype=bg.F64)
    opts = bg.gmres_opts(m=32, tol=args.tol, max_iter=10000, reorth_eta=0.1, orth={"cgs": bg.ORTH_CGS_ADAPTIVE, "cgs2": bg.ORTH_CGS2, "mgs": bg.ORTH_MGS}[args.orth])
    import scipy.sparse as sp

    # sources: column 0 = the supply / pad network (constant), column 1 = the switching loads, swinging +-30 % as a
    # PWL table sampled from a sine (10 breakpoints per 200 steps); B carries b0 split into its positive (pads) and
    # negative (loads) entries
    pads = np.flatnonzero(b0 > 0); loads = np.flatnonzero(b0 < 0)
    B = sp.csr_matrix((np.concatenate([b0[pads], b0[loads]]), (np.concatenate([pads, loads]), np.concatenate([np.zeros(len(pads), int), np.ones(len(loads), int)]))), shape=(n, 2))
    B.sort_indices()
    tk = np.arange(0, args.steps + 21, 20, dtype=np.float64)
    pwl = [(np.array([0.0]), np.array([1.0])), (tk, 1.0 + 0.3 * np.sin(2 * np.pi * tk / 200.0))]
    Ch = sp.diags(c_h).tocsr()
    T = bg.Transient(ctx, S, A, M, (Ch.indptr, Ch.indices, Ch.data), (B.indptr, B.indices, B.data), pwl, probes=[0, n // 2, n - 1])
    dx = torch.zeros(n, dtype=torch.float64, device="cuda")
    d_b0 = torch.from_numpy(b0).cuda()
    torch.cuda.synchronize()
    # DC operating point (x = 0 start), then the time loop (unit time step: the table is in steps)
    r = S.solve(A, M, d_b0.data_ptr(), dx.data_ptr(), opts)
    dc_iters, dc_ms = int(r.iters), r.solve_ms
    t_wall = time.perf_counter()
    rc, st, probe = T.run(dx.data_ptr(), 0.0, 1.0, args.steps, opts)
    wall = time.perf_counter() - t_wall
    if rc != 0:
        raise SystemExit(f"{st.not_converged} steps did not converge")
    ms = st.device_ms
    iters = [st.iters_total / args.steps]
    iters_max = st.iters_max
    line = {"config": f"BASELINE configs[3]: pg2d {nx}x{nx} transient, {args.steps} backward-Euler steps, ILU(0) factor reused, warm-started GMRES(32), whole loop on the device (bg_transient_run)",
            "n": n, "steps": args.steps, "ms_per_step": ms / args.steps, "total_ms": ms, "wall_s": wall, "iters_total": int(st.iters_total), "iters_per_step_mean": float(iters[0]),
            "iters_per_step_max": int(iters_max), "dc_iters": dc_iters, "dc_ms": dc_ms, "last_resid": float(st.last_resid),
            "probe_last_step": [float(p) for p in probe[-1]], "kernel_launches": int(st.launches),
            "us_per_iteration": 1e3 * ms / max(int(st.iters_total), 1), "setup_s": setup}
    line["orth"] = args.orth
    print(json.dumps(line), flush=True)
    T.close(); S.close(); M.close(); A.close(); ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--orth", default="cgs", choices=["cgs", "cgs2", "mgs"], help="cgs: fused adaptive CGS step (performance mode); mgs: the reference's strict order (the ABI default)")
    ap.add_argument("--which", default="thermal,transient")
    ap.add_argument("--thermal-n", type=int, default=171)
    ap.add_argument("--thermal-tile", type=int, default=10)
    ap.add_argument("--transient-nx", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tol", type=float, default=1e-8)
    args = ap.parse_args()
    import torch

    import b200gmres as bg

    for w in args.which.split(","):
        {"thermal": thermal, "transient": transient}[w](args, bg, torch)


if __name__ == "__main__":
    main()
