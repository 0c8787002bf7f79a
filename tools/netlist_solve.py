#!/usr/bin/env python
"""A SPICE-subset power-grid netlist end to end on the device: parse (bg_netlist_*), DC operating point with
GMRES(m) + ILU(0), and -- when the netlist has a .tran line -- the backward-Euler loop of bg_transient_* with the
.print nodes sampled.  The role of parser_wrapper + mna_solve_gpu_gmres in the reference
(src/etbr_wrapper.cpp:46-171, src/mna_solve_gpu_gmres.cpp:437-647).

    python tools/netlist_solve.py grid.sp [--mna] [--dims NX NY] [--out ports.txt] [--tol 1e-8]
    python tools/netlist_solve.py --generate 300 300 --nsteps 50 --load pwl /tmp/grid.sp     # write a synthetic grid

By default grounded voltage sources are eliminated (BG_NETLIST_REDUCED) so that ILU(0) applies; --mna keeps the
reference's full MNA system (zero diagonals on the source rows: ILU(0) stops with BG_ERR_ZERO_PIVOT unless the
matrix happens to allow it -- the reference needs the pivoting ILU++ there).  --dims NX NY declares the unknowns
an NX x NY grid in row-major order (as --generate writes them) and selects the block-Jacobi lattice schedule.
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import numpy as np


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("netlist")
    ap.add_argument("--generate", nargs=2, type=int, metavar=("NX", "NY"))
    ap.add_argument("--nsteps", type=int, default=0)
    ap.add_argument("--load", default="dc", choices=["dc", "pwl", "pulse"])
    ap.add_argument("--mna", action="store_true")
    ap.add_argument("--dims", nargs=2, type=int, metavar=("NX", "NY"))
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--restart", type=int, default=32)
    ap.add_argument("--out")
    args = ap.parse_args()
    import b200gmres as bg
    from b200gmres import synth
    if args.generate:
        nx, ny = args.generate
        with open(args.netlist, "w") as f:
            f.write(synth.pg2d_netlist(nx, ny, seed=synth.SEED_BASE + 2, nsteps=args.nsteps, load=args.load))
        print(f"wrote {args.netlist}: {nx} x {ny} grid")
        return 0
    import torch
    t0 = time.perf_counter()
    N = bg.Netlist(args.netlist, formulation=bg.NETLIST_MNA if args.mna else bg.NETLIST_REDUCED)
    i = N.info
    print(f"parsed in {time.perf_counter() - t0:.2f} s: {i.n_nodes} nodes, {i.n_resistors} R, {i.n_capacitors} C, {i.n_inductors} L, "
          f"{i.n_vsources} V, {i.n_isources} I; {i.size} unknowns ({i.n_eliminated} nodes eliminated), nnz(G) = {i.nnz_g}")
    n = i.size
    ctx = bg.Context(0)
    if args.dims:
        nx, ny = args.dims
        assert nx * ny == n, "--dims must describe all unknowns"
        popts = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=1024, block_method=bg.PART_GRID2D, dims=[nx, ny, 1, 32, 32, 1])
    else:
        popts = bg.precond_opts(trsv_schedule=bg.TRSV_LEVEL)
    opts = bg.gmres_opts(m=args.restart, tol=args.tol, max_iter=20000, orth=bg.ORTH_CGS_ADAPTIVE if args.restart <= 32 else bg.ORTH_MGS, reorth_eta=0.1)
    import scipy.sparse as sp
    G = sp.csr_matrix((N.G[2], N.G[1], N.G[0]), shape=(n, n))
    B = sp.csr_matrix((N.B[2], N.B[1], N.B[0]), shape=(n, i.n_sources))
    # DC operating point: G x = B u(0)
    A = bg.Matrix(ctx, *N.G, dtype=bg.F64)
    M = bg.Precond.ilu0(ctx, *N.G, dtype=bg.F64, opts=popts)
    S = bg.Solver(ctx, n, args.restart, dtype=bg.F64)
    b = B @ N.u(0.0)
    db = torch.from_numpy(b).cuda()
    dx = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), opts)
    ctx.sync()
    x = dx.cpu().numpy()
    true = np.linalg.norm(b - G @ x) / max(np.linalg.norm(b), 1e-300)
    print(f"DC: flag {r.flag}, {r.iters} iterations, {r.solve_ms:.2f} ms on the device, true relative residual {true:.2e}")
    ports = [(N.row_name(p) if p >= 0 else "(fixed)", p) for p in N.ports]
    for name, p in ports[:8]:
        if p >= 0:
            print(f"   v({name}) = {x[p]:.9f}")
    if i.dc or i.tstep <= 0:
        if args.out:
            np.savetxt(args.out, x)
        return 0 if r.flag == 0 else 1
    # transient: (G + C/h) x_i = B u(t_i) + (C/h) x_{i-1}
    h = i.tstep
    nsteps = int(round(i.tstop / h))
    Ch = sp.csr_matrix((N.C[2] / h, N.C[1], N.C[0]), shape=(n, n))
    At = (G + Ch).tocsr(); At.sort_indices()
    A2 = bg.Matrix(ctx, At.indptr, At.indices, At.data, dtype=bg.F64)
    M2 = bg.Precond.ilu0(ctx, At.indptr, At.indices, At.data, dtype=bg.F64, opts=popts)
    probes = [p for p in N.ports if p >= 0]
    T = bg.Transient(ctx, S, A2, M2, (Ch.indptr, Ch.indices, Ch.data), N.B, N.transient_pwl, probes)
    if i.n_pulse:
        T.set_pulse(N.pulse_src, N.pulse_time, N.pulse_val)
    rc, st, probe = T.run(dx.data_ptr(), 0.0, h, nsteps, opts)
    ctx.sync()
    print(f"transient: {st.steps} steps of {h:g} s, {st.iters_total} iterations in total (max {st.iters_max} per step), "
          f"{st.not_converged} not converged, {st.device_ms:.1f} ms on the device ({st.device_ms / max(st.steps, 1):.3f} ms per step)")
    if args.out and probe is not None:
        np.savetxt(args.out, np.column_stack([h * np.arange(1, nsteps + 1), probe]), header="t " + " ".join(nm for nm, p in ports if p >= 0))
    return 0 if rc == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
