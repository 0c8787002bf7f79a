#!/usr/bin/env python
"""Per-kernel timings of the hot path on one B200 (development aid; bench.py is the contract).

Times, with CUDA events on the library's own stream after warm-up:
  * y = A x                       (spmv_sell_* / spmv_csr_vector)
  * z = U^-1 L^-1 r               (trsv_block_kernel x2 or the level schedule)
  * one fused orthogonalisation   (dots_kernel + update_kernel) for several basis sizes j
  * one full solve with per-phase event pairs (bg_solver_profile)
and prints achieved GB/s against the algorithmic bytes of SURVEY.md 8(d).

    python tools/kernel_bench.py --nx 1304 --tiles 16,32 --factor-dtypes f64,f32
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1304)
    ap.add_argument("--tiles", default="32")
    ap.add_argument("--factor-dtypes", default="f64")
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--orths", default="cgs")
    ap.add_argument("--precond", default="ilu0")
    ap.add_argument("--schedules", default="block")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--js", default="1,4,8,16,24,32")
    ap.add_argument("--tma", default="1")
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--etas", default="0.1")
    ap.add_argument("--chunks", default="4096")
    ap.add_argument("--stages", default="2")
    ap.add_argument("--warps", default="4")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch

    import b200gmres as bg
    from b200gmres import synth

    peak = 6541.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    dtc = bg.F32 if args.dtype == "f32" else bg.F64
    npdt = np.float32 if dtc == bg.F32 else np.float64
    s = 4 if dtc == bg.F32 else 8
    rp, ci, v, b = synth.pg2d(args.nx, args.nx, seed=synth.SEED_BASE + 2)
    n = len(rp) - 1
    ctx = bg.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    A = bg.Matrix(ctx, rp, ci, v, dtype=dtc)
    S = bg.Solver(ctx, n, 32, dtype=dtc)
    db = torch.from_numpy(b.astype(npdt)).cuda()
    dx = torch.zeros(n, dtype=db.dtype, device="cuda")
    dy = torch.zeros(n, dtype=db.dtype, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    rows = []

    def timeit(fn, reps, flush_l2=False):
        for _ in range(3):
            fn()
        ctx.sync()
        tot = 0.0
        if flush_l2:
            for _ in range(reps):
                with torch.cuda.stream(stream):
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / reps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def report(name, ms, nbytes, **kw):
        gbs = nbytes / (ms * 1e-3) / 1e9
        row = dict(kernel=name, us=round(ms * 1e3, 2), MB=round(nbytes / 1e6, 1), GBps=round(gbs, 1), frac=round(gbs / peak, 3), **kw)
        rows.append(row)
        print(json.dumps(row), flush=True)

    info = A.info()
    for fl in (False, True):
        ms = timeit(lambda: A.spmv(db.data_ptr(), dy.data_ptr()), args.reps, fl)
        report("spmv", ms, info.spmv_bytes, l2_flushed=fl, fmt=int(info.format))

    for j in [int(t) for t in args.js.split(",")]:
        ms = timeit(lambda: S.bench_orth(j), args.reps)
        report("orth(dots+update)", ms, (2 * j + 5) * s * n, j=j)

    for sched in args.schedules.split(","):
        for tile in [int(t) for t in args.tiles.split(",")]:
            for fd in args.factor_dtypes.split(","):
                import itertools
                for tma, chunk, stages, warps in itertools.product([int(t) for t in args.tma.split(",")], [int(t) for t in args.chunks.split(",")],
                                                                   [int(t) for t in args.stages.split(",")], [int(t) for t in args.warps.split(",")]):
                    fdt = bg.F32 if fd == "f32" else bg.F64
                    if sched == "block":
                        o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=tile * tile, factor_dtype=fdt, use_tma=tma,
                                            chunk_bytes=chunk, stages=stages, warps_per_cta=warps,
                                            block_method=bg.PART_GRID2D, dims=[args.nx, args.nx, 1, tile, tile, 1])
                    else:
                        o = bg.precond_opts(trsv_schedule=bg.TRSV_LEVEL, factor_dtype=fdt)
                    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=dtc, opts=o, ilut=(args.precond == "ilut"))
                    pi = M.info()
                    tag = dict(schedule=sched, tile=tile, factor=fd, chunk=chunk, stages=stages, warps=warps, blocks=pi.nblocks, levels=(pi.levelsL, pi.levelsU))
                    reps = args.reps if sched == "block" else 3
                    for fl in (False, True):
                        ms = timeit(lambda: M.apply(bg.PC_FULL, db.data_ptr(), dy.data_ptr()), reps, fl)
                        report("trsv(L+U)", ms, pi.trsv_bytes, l2_flushed=fl, stream_MB=round(pi.stream_bytes / 1e6, 1), **tag)
                    for orth, eta in itertools.product(args.orths.split(","), [float(t) for t in args.etas.split(",")]):
                        if orth != "cgs" and eta != float(args.etas.split(",")[0]):
                            continue
                        oc = {"mgs": bg.ORTH_MGS, "cgs": bg.ORTH_CGS_ADAPTIVE, "cgs2": bg.ORTH_CGS2}[orth]
                        opts = bg.gmres_opts(m=32, tol=args.tol, max_iter=3000, orth=oc, use_graph=args.graph, reorth_eta=eta)
                        for _ in range(2):
                            with torch.cuda.stream(stream):
                                dx.zero_()
                            r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), opts)
                        with torch.cuda.stream(stream):
                            dx.zero_()
                        rp_, st = S.profile(A, M, db.data_ptr(), dx.data_ptr(), opts)
                        tot = st.spmv_bytes + st.trsv_bytes + st.orth_bytes + st.other_bytes
                        row = dict(kernel="solve", orth=orth, eta=eta, resid=float(r.resid), iters=r.iters, reorths=r.reorths, solve_ms=round(r.solve_ms, 3),
                                   us_per_iter=round(1e3 * r.solve_ms / max(r.iters_done, 1), 1), launches=int(r.launches),
                                   GBps=round(tot / (r.solve_ms * 1e-3) / 1e9, 1), frac=round(tot / (r.solve_ms * 1e-3) / 1e9 / peak, 3),
                                   profiled_ms=dict(spmv=round(st.spmv_ms, 3), trsv=round(st.trsv_ms, 3), orth=round(st.orth_ms, 3),
                                                    other=round(st.other_ms, 3)), **tag)
                        rows.append(row)
                        print(json.dumps(row), flush=True)
                    M.close()
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)
    S.close(); A.close(); ctx.close()


if __name__ == "__main__":
    main()
