"""Device-resident backward-Euler loop (bg_transient_*) against a host time-stepping with direct solves:
the reference's loop (src/mna_solve_gpu_gmres.cpp:564-647) computes rhs = B u(t_i) + (C/h) x_{i-1} on the CPU, solves
with the factor built once, warm-starts from x_{i-1} and samples the ports."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("dtc", [1, 0])
def test_transient_matches_host_time_stepping(bg, ctx, dtc):
    from b200gmres import synth

    nx, ny = 40, 30
    rp, ci, v, b0 = synth.pg2d(nx, ny, seed=31, c_over_h=0.2)
    _, _, v0, _ = synth.pg2d(nx, ny, seed=31, c_over_h=0.0)
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    G = sp.csr_matrix((v0, ci, rp), shape=(n, n))
    Ch = sp.csr_matrix(A - G)
    Ch.eliminate_zeros(); Ch.sort_indices()
    # sources: 3 PWL current sources driving random node subsets + 1 constant supply carrying the DC right-hand side
    rng = np.random.default_rng(5)
    nsrc = 4
    cols, rows, vals = [], [], []
    for k in range(3):
        nodes = rng.choice(n, size=25, replace=False)
        rows += list(nodes); cols += [k] * len(nodes); vals += list(-rng.uniform(0.5e-3, 2e-3, size=len(nodes)))
    nzb = np.flatnonzero(b0)
    rows += list(nzb); cols += [3] * len(nzb); vals += list(b0[nzb])
    B = sp.csr_matrix((vals, (rows, cols)), shape=(n, nsrc)); B.sort_indices()
    pwl = [(np.array([0.0, 2.0, 5.0, 9.0]), np.array([0.0, 1.0, 0.2, 1.5])),
           (np.array([1.0, 3.0, 4.0]), np.array([1.0, 0.0, 2.0])),
           (np.array([0.0, 20.0]), np.array([0.5, 2.5])),
           (np.array([0.0]), np.array([1.0]))]
    probes = [0, n // 3, n - 1]
    h, nsteps = 0.5, 24
    npdt, tdt = (np.float64, torch.float64) if dtc == 1 else (np.float32, torch.float32)
    tol = 1e-10 if dtc == 1 else 1e-6

    def u_of(t):
        return np.array([np.interp(t, tt, vv) for tt, vv in pwl])

    lu = spla.splu(A.tocsc())
    x_ref = lu.solve(B @ u_of(0.0))  # operating point at t0
    x0 = x_ref.copy()
    probe_ref = np.zeros((nsteps, len(probes)))
    for i in range(1, nsteps + 1):
        x_ref = lu.solve(B @ u_of(h * i) + Ch @ x_ref)
        probe_ref[i - 1] = x_ref[probes]

    Am = bg.Matrix(ctx, rp, ci, v, dtype=dtc)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=dtc)
    S = bg.Solver(ctx, n, 32, dtype=dtc)
    T = bg.Transient(ctx, S, Am, M, (Ch.indptr, Ch.indices, Ch.data), (B.indptr, B.indices, B.data), pwl, probes)
    dx = torch.from_numpy(x0.astype(npdt)).cuda()
    torch.cuda.synchronize()
    rc, st, probe = T.run(dx.data_ptr(), 0.0, h, nsteps, bg.gmres_opts(m=32, tol=tol, max_iter=500, orth=bg.ORTH_MGS))
    ctx.sync()
    x = dx.cpu().numpy().astype(np.float64)
    assert rc == 0 and st.steps == nsteps and st.not_converged == 0
    lim = 1e-7 if dtc == 1 else 2e-3
    assert np.linalg.norm(x - x_ref) <= lim * np.linalg.norm(x_ref)
    assert np.abs(probe - probe_ref).max() <= lim * np.abs(probe_ref).max()
    # warm start pays: far fewer iterations per step than a cold DC solve needs
    dc = torch.zeros(n, dtype=tdt, device="cuda")
    db = torch.from_numpy((B @ u_of(0.0)).astype(npdt)).cuda()
    torch.cuda.synchronize()
    r_dc = S.solve(Am, M, db.data_ptr(), dc.data_ptr(), bg.gmres_opts(m=32, tol=tol, max_iter=500, orth=bg.ORTH_MGS))
    assert st.iters_total / nsteps < r_dc.iters
    T.close(); S.close(); M.close(); Am.close()


def pulse_value(t, td, tr, tf, tw, tp, vlo, vhi):
    """gen_PULSEut_kernel (src/kernels.cu:224-246), restated for one time"""
    mt = t - np.floor(t / tp) * tp
    if mt < td:
        return vlo
    if mt < td + tr:
        return vlo + (mt - td) * (vhi - vlo) / tr
    if mt < td + tr + tw:
        return vhi
    if mt < td + tr + tw + tf:
        return vhi - (mt - td - tr - tw) * (vhi - vlo) / tf
    return vlo


def test_transient_pulse_sources(bg, ctx):
    """PULSE current sources (periodic trapezoids, src/kernels.cu:224-246) next to a piecewise-linear one, in the
    device-resident loop, against host time stepping with direct solves."""
    from b200gmres import synth

    nx, ny = 32, 24
    rp, ci, v, b0 = synth.pg2d(nx, ny, seed=47, c_over_h=0.3)
    _, _, v0, _ = synth.pg2d(nx, ny, seed=47, c_over_h=0.0)
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    Ch = sp.csr_matrix(A - sp.csr_matrix((v0, ci, rp), shape=(n, n)))
    Ch.eliminate_zeros(); Ch.sort_indices()
    rng = np.random.default_rng(9)
    nsrc = 4
    rows, cols, vals = [], [], []
    for k in range(3):
        nodes = rng.choice(n, size=20, replace=False)
        rows += list(nodes); cols += [k] * len(nodes); vals += list(-rng.uniform(0.5e-3, 2e-3, size=len(nodes)))
    nzb = np.flatnonzero(b0)
    rows += list(nzb); cols += [3] * len(nzb); vals += list(b0[nzb])
    B = sp.csr_matrix((vals, (rows, cols)), shape=(n, nsrc)); B.sort_indices()
    pwl = [(np.array([0.0]), np.array([0.0])), (np.array([0.0, 4.0, 6.0]), np.array([0.0, 1.0, 0.5])),
           (np.array([0.0]), np.array([0.0])), (np.array([0.0]), np.array([1.0]))]
    pulses = {0: ((0.7, 0.5, 0.9, 1.3, 4.0), (0.1, 1.7)), 2: ((0.0, 0.25, 0.25, 0.5, 1.5), (-0.4, 0.8))}
    h, nsteps = 0.3, 40

    def u_of(t):
        u = np.array([np.interp(t, tt, vv) for tt, vv in pwl])
        for k, (tm, vl) in pulses.items():
            u[k] = pulse_value(t, *tm, *vl)
        return u

    lu = spla.splu(A.tocsc())
    x_ref = lu.solve(B @ u_of(0.0))
    x0 = x_ref.copy()
    probes = [1, n // 2]
    probe_ref = np.zeros((nsteps, len(probes)))
    for i in range(1, nsteps + 1):
        x_ref = lu.solve(B @ u_of(h * i) + Ch @ x_ref)
        probe_ref[i - 1] = x_ref[probes]
    Am = bg.Matrix(ctx, rp, ci, v, dtype=1)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1)
    S = bg.Solver(ctx, n, 32, dtype=1)
    T = bg.Transient(ctx, S, Am, M, (Ch.indptr, Ch.indices, Ch.data), (B.indptr, B.indices, B.data), pwl, probes)
    T.set_pulse(list(pulses), [pulses[k][0] for k in pulses], [pulses[k][1] for k in pulses])
    with pytest.raises(bg.BgError):
        T.set_pulse([nsrc], [(0, 1, 1, 1, 1)], [(0, 1)])       # source index out of range
    with pytest.raises(bg.BgError):
        T.set_pulse([0], [(0, 1, 1, 1, 0.0)], [(0, 1)])        # period must be positive
    dx = torch.from_numpy(x0).cuda()
    torch.cuda.synchronize()
    rc, st, probe = T.run(dx.data_ptr(), 0.0, h, nsteps, bg.gmres_opts(m=32, tol=1e-10, max_iter=500, orth=bg.ORTH_MGS))
    ctx.sync()
    x = dx.cpu().numpy()
    assert rc == 0 and st.not_converged == 0
    assert np.linalg.norm(x - x_ref) <= 1e-7 * np.linalg.norm(x_ref)
    assert np.abs(probe - probe_ref).max() <= 1e-7 * np.abs(probe_ref).max()
    T.close(); S.close(); M.close(); Am.close()


@pytest.mark.parametrize("load", ["pwl", "pulse"])
def test_netlist_to_device_transient(bg, ctx, load):
    """netlist text -> bg_netlist_* (reduced formulation) -> DC operating point + device-resident transient with the
    .print nodes sampled, against host backward Euler with direct solves on the FULL MNA system of the same netlist
    (voltage-source rows kept): the two formulations and the two code paths must describe the same circuit."""
    from b200gmres import synth

    nx, ny, nsteps = 48, 40, 30
    text = synth.pg2d_netlist(nx, ny, seed=synth.SEED_BASE + 9, pad_pitch=8, nsteps=nsteps, load=load, ports=((0, 0), (24, 20), (47, 39)))
    F = bg.Netlist(text, text=True)                                       # full MNA: the host reference
    R = bg.Netlist(text, formulation=bg.NETLIST_REDUCED, text=True)     # what runs on the device
    h = R.info.tstep
    assert R.info.size == nx * ny and abs(R.info.tstop - nsteps * h) < 1e-25

    def csr(t, shape):
        return sp.csr_matrix((t[2], t[1], t[0]), shape=shape)

    nf, nr, ns = F.info.size, R.info.size, R.info.n_sources
    Gf, Cf, Bf = csr(F.G, (nf, nf)), csr(F.C, (nf, nf)), csr(F.B, (nf, ns))

    def u_of(t):
        u = F.u(t)
        for k, (tm, pv) in zip(F.pulse_src, zip(F.pulse_time, F.pulse_val)):
            u[k] = pulse_value(t, *tm, *pv)
        return u

    x_ref = spla.splu(Gf.tocsc()).solve(Bf @ u_of(0.0))
    x0_full = x_ref.copy()
    lu = spla.splu((Gf + Cf / h).tocsc())
    probe_ref = np.zeros((nsteps, len(F.ports)))
    for i in range(1, nsteps + 1):
        x_ref = lu.solve(Bf @ u_of(h * i) + (Cf / h) @ x_ref)
        probe_ref[i - 1] = x_ref[F.ports]

    Gr, Cr = csr(R.G, (nr, nr)), csr(R.C, (nr, nr))
    Ch = sp.csr_matrix(Cr / h); Ch.sort_indices()
    At = sp.csr_matrix(Gr + Ch); At.sort_indices()
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=1024, block_method=bg.PART_GRID2D, dims=[nx, ny, 1, 32, 32, 1])
    opts = bg.gmres_opts(m=32, tol=1e-11, max_iter=2000, orth=bg.ORTH_CGS_ADAPTIVE, reorth_eta=0.1)
    S = bg.Solver(ctx, nr, 32, dtype=bg.F64)
    # DC operating point on the device
    A0 = bg.Matrix(ctx, *R.G, dtype=bg.F64)
    M0 = bg.Precond.ilu0(ctx, *R.G, dtype=bg.F64, opts=o)
    db = torch.from_numpy(csr(R.B, (nr, ns)) @ u_of(0.0)).cuda()
    dx = torch.zeros(nr, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    r = S.solve(A0, M0, db.data_ptr(), dx.data_ptr(), opts)
    ctx.sync()
    assert r.flag == 0
    nn = F.info.n_nodes
    full0 = R.expand(dx.cpu().numpy(), 0.0)
    assert np.abs(full0[:nn] - x0_full[:nn]).max() <= 1e-8 * np.abs(x0_full[:nn]).max()
    # transient on the device, started from the device's own operating point
    A1 = bg.Matrix(ctx, At.indptr, At.indices, At.data, dtype=bg.F64)
    M1 = bg.Precond.ilu0(ctx, At.indptr, At.indices, At.data, dtype=bg.F64, opts=o)
    T = bg.Transient(ctx, S, A1, M1, (Ch.indptr, Ch.indices, Ch.data), R.B, R.transient_pwl, list(R.ports))
    if R.info.n_pulse:
        T.set_pulse(R.pulse_src, R.pulse_time, R.pulse_val)
    rc, st, probe = T.run(dx.data_ptr(), 0.0, h, nsteps, opts)
    ctx.sync()
    assert rc == 0 and st.steps == nsteps and st.not_converged == 0
    full = R.expand(dx.cpu().numpy(), h * nsteps)
    assert np.abs(full[:nn] - x_ref[:nn]).max() <= 1e-7 * np.abs(x_ref[:nn]).max()
    assert np.abs(probe - probe_ref).max() <= 1e-7 * np.abs(probe_ref).max()
    assert np.ptp(probe_ref[:, 1]) > 1e-6        # the loads do move the grid: the comparison is not of a constant
    T.close(); S.close(); M1.close(); A1.close(); M0.close(); A0.close()
