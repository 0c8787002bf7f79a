"""GPU parity (B200): every hot-path kernel and the full GMRES(m) solve, called through the C ABI,
against the CPU oracle on the same seeded inputs and against the committed golden fixtures.

Tolerances (BASELINE.json north_star): integer artefacts (level sets) bit-exact; FP64 final
residual within 1e-10 (absolute difference of the relative residuals) and solution within 1e-8
relative; FP32 modes are compared at FP32 round-off level (reduction order differs by design)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

X_RTOL_F64 = 1e-8      # ||x_gpu - x_oracle|| / ||x_oracle||
RESID_ATOL_F64 = 1e-10  # |resid_gpu - resid_oracle| on relative residuals


def dev(a, dt):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    torch.cuda.synchronize()
    return t


def host(ctx, t):
    ctx.sync()
    return t.cpu().numpy()


def np_dt(code):
    return np.float32 if code == 0 else np.float64


def gen_cases():
    from b200gmres import synth

    return {
        "pg2d_61x47": synth.pg2d(61, 47, seed=101),
        "th3d_13x11x9": synth.th3d(13, 11, 9, seed=103),
        "rand_2000": synth.random_sparse(2000, 7, seed=107),
        "rand_long": synth.random_sparse(700, 40, seed=109),
    }


CASES = gen_cases()


# ------------------------------------------------------------------------------- SpMV
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("dtc", [0, 1])
@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
def test_spmv_matches_oracle(bg, ctx, oracle, name, dtc, fmt):
    rp, ci, v, b = CASES[name]
    dt = np_dt(dtc)
    n = len(rp) - 1
    A = bg.Matrix(ctx, rp, ci, v, dtype=dtc, fmt=fmt)
    x = np.cos(np.arange(n) * 0.37) + 0.5
    dx, dy = dev(x, dt), dev(np.zeros(n), dt)
    A.spmv(dx.data_ptr(), dy.data_ptr())
    y = host(ctx, dy)
    want = oracle.spmv(rp, ci, v, x, "f32" if dtc == 0 else "f64").astype(np.float64)
    # the reference's own acceptance test: ||ref - gpu|| / ||ref|| < 1e-6 (src_thermal/main1.cu:292-308)
    assert np.linalg.norm(y - want) / np.linalg.norm(want) < (1e-6 if dtc == 0 else 1e-14)
    # fused residual form  rr = b - A x  (gmres.cu:2302-2305)
    db, dr = dev(b, dt), dev(np.zeros(n), dt)
    A.spmv_axpby(-1.0, dx.data_ptr(), 1.0, db.data_ptr(), dr.data_ptr())
    r = host(ctx, dr)
    assert np.linalg.norm(r - (b.astype(dt) - want)) / np.linalg.norm(want) < (2e-6 if dtc == 0 else 1e-14)
    A.close()


def test_spmv_edge_rows(bg, ctx):
    """empty rows, a single dense row, n not a multiple of the slice height"""
    n = 131
    rows = [0] * n + list(range(2, n, 3))
    cols = list(range(n)) + [0] * len(range(2, n, 3))
    vals = np.arange(1, len(rows) + 1, dtype=np.float64)
    M = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    M.sort_indices()
    x = np.linspace(-1, 1, n)
    for fmt in (1, 2, 3):
        for dtc in (0, 1):
            A = bg.Matrix(ctx, M.indptr, M.indices, M.data, dtype=dtc, fmt=fmt)
            dx, dy = dev(x, np_dt(dtc)), dev(np.full(n, 7.0), np_dt(dtc))
            A.spmv(dx.data_ptr(), dy.data_ptr())
            y = host(ctx, dy)
            assert np.allclose(y, M @ x, rtol=1e-5 if dtc == 0 else 1e-13, atol=1e-5 if dtc == 0 else 1e-13)
            A.close()


def test_inspector_picks_format(bg, ctx):
    rp, ci, v, _ = CASES["pg2d_61x47"]
    A = bg.Matrix(ctx, rp, ci, v)
    i = A.info()
    assert i.format == bg.SPMV_SELL and i.max_row_nnz == 5 and i.hist[3] > 0
    assert i.spmv_bytes == 12 * i.nnz + 4 * (i.n + 1) + 16 * i.n
    A.close()
    rp, ci, v, _ = CASES["rand_long"]
    A = bg.Matrix(ctx, rp, ci, v)
    assert A.info().format in (bg.SPMV_CSR_VECTOR, bg.SPMV_CSR_WARP)
    A.close()


# ------------------------------------------------------------------------------- preconditioner
def oracle_ilu0_pc(oracle, bg, rp, ci, v, mode, blk=None):
    dt = np.float32 if mode == "f32" else np.float64
    f = bg.host_ilu0(rp, ci, v.astype(dt), blk)  # bit-exact with the oracle's (test_host_logic)
    n = len(rp) - 1
    return f, oracle.Precond(oracle.PC_ILU0, n, mode, **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})


@pytest.mark.parametrize("name", ["pg2d_61x47", "th3d_13x11x9", "rand_2000"])
@pytest.mark.parametrize("dtc", [0, 1])
@pytest.mark.parametrize("lattice", [1, 0])
def test_ilu0_level_schedule_matches_oracle(bg, ctx, oracle, name, dtc, lattice):
    """Reference operator (global factors): the wavefront-strip kernel where the factors are a 5-point grid
    (lattice=1 on pg2d), the sync-free slice kernel everywhere else."""
    rp, ci, v, b = CASES[name]
    mode = "f32" if dtc == 0 else "f64"
    dt = np_dt(dtc)
    n = len(rp) - 1
    f, pc = oracle_ilu0_pc(oracle, bg, rp, ci, v, mode)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=dtc, opts=bg.precond_opts(lattice=lattice))
    assert M.info().lattice == (1 if lattice and name.startswith("pg2d") else 0)
    # level sets the device schedule was built from == getLevel (bit-exact integer artefact)
    assert np.array_equal(M.levels(True), oracle.get_level(1, f["lp"], f["li"], f["lv"], mode))
    assert np.array_equal(M.levels(False), oracle.get_level(0, f["up"], f["ui"], f["uv"], mode))
    info = M.info()
    assert (info.nnzL, info.nnzU) == (len(f["lv"]), len(f["uv"]))
    din, dout = dev(b, dt), dev(np.zeros(n), dt)
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    got = host(ctx, dout)
    want = pc.apply(0, b)
    # same operations in the same order per row: FP64 agrees to the last bits, FP32 to a few ulp (FMA); the
    # strip kernel solves the row-scaled unit system D^-1 U, which moves the last bit or two
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < (5e-6 if dtc == 0 else 1e-14)
    # a second application must not see the first one's packets (sequence numbers)
    dout.zero_()
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    assert np.array_equal(host(ctx, dout), got)
    M.close()


def test_level_schedule_with_explicit_tiny_entries(bg, ctx):
    """caller-supplied factors holding explicit |v| < 1e-9 entries.  getLevel stops at / skips them
    (the parity artefact keeps doing so), the device schedule follows the structural dependencies, so every
    stored entry is applied in order."""
    rng = np.random.default_rng(5)
    n = 900
    L = sp.tril(sp.random(n, n, density=0.004, random_state=rng, format="csr"), -1).tolil()
    U = sp.triu(sp.random(n, n, density=0.004, random_state=rng, format="csr"), 1).tolil()
    for i in range(1, n):
        L[i, i - 1] = 1e-12 if i % 3 == 0 else 0.4   # the FIRST entry of many rows is tiny
    for i in range(n - 1):
        U[i, i + 1] = 1e-12 if i % 4 == 0 else -0.3
    L = (L.tocsr() + sp.eye(n)).tocsr(); U = (U.tocsr() + 2.5 * sp.eye(n)).tocsr()
    L.sort_indices(); U.sort_indices()
    b = rng.uniform(-1, 1, n)
    M = bg.Precond.lu(ctx, L.indptr, L.indices, L.data, U.indptr, U.indices, U.data, dtype=1)
    din, dout = dev(b, np.float64), dev(np.zeros(n), np.float64)
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    got = host(ctx, dout)
    Lu = sp.tril(L, -1) + sp.eye(n)  # unit lower: the stored diagonal is ignored (DIAG_TYPE_UNIT)
    want = sp.linalg.spsolve_triangular(U.tocsr(), sp.linalg.spsolve_triangular(Lu.tocsr(), b, lower=True), lower=False)
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)
    assert np.array_equal(M.levels(True), bg.host_trsv_levels(1, L.indptr, L.indices, L.data))  # parity API unchanged
    M.close()


@pytest.mark.parametrize("name,method,dims", [
    ("pg2d_61x47", 2, [61, 47, 1, 16, 16, 1]),
    ("pg2d_61x47", 1, None),
    ("th3d_13x11x9", 3, [13, 11, 9, 6, 6, 5]),
    ("rand_2000", 0, None),
])
@pytest.mark.parametrize("dtc,fdt", [(1, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("tma", [1, 0])
def test_block_jacobi_schedule_matches_oracle(bg, ctx, oracle, name, method, dims, dtc, fdt, tma):
    rp, ci, v, b = CASES[name]
    dt = np_dt(dtc)
    n = len(rp) - 1
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=method, block_rows=256, factor_dtype=fdt, use_tma=tma,
                        **({"dims": dims} if dims else {}))
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=dtc, opts=o)
    info = M.info()
    assert info.nblocks > 1 and info.max_block_rows <= 1024
    # rebuild the same blocks on the host to form the oracle's block-Jacobi ILU(0)
    nb, blk = bg.host_block_partition(method, rp, ci, max(1, (n + 255) // 256), dims)
    assert nb == info.nblocks
    _, pc = oracle_ilu0_pc(oracle, bg, rp, ci, v, "f32" if dtc == 0 else "f64", blk)
    din, dout = dev(b, dt), dev(np.zeros(n), dt)
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    got = host(ctx, dout)
    want = pc.apply(0, b)
    tol = 1e-13 if (dtc, fdt) == (1, 1) else 5e-6
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol
    M.close()


@pytest.mark.parametrize("nx,ny,tx,ty", [(61, 47, 16, 16), (96, 80, 32, 32), (70, 33, 32, 8), (40, 100, 7, 32)])
@pytest.mark.parametrize("dtc", [0, 1])
def test_device_factorisation_equals_host(bg, ctx, nx, ny, tx, ty, dtc):
    """block-Jacobi ILU(0) of grid tiles factored AND packed on the device (factor.cu) is bit-identical
    with the host pipeline (bg_host_ilu0 = leftILU semantics, then the stream packer): same operations, same order,
    no FMA contraction.  Partial tiles at the right / top edge, tiles wider than tall and the other way round."""
    from b200gmres import synth

    rp, ci, v, b = synth.pg2d(nx, ny, seed=311)
    n = len(rp) - 1
    dt = np_dt(dtc)
    outs, infos = [], []
    for dev_factor in (1, 0):
        o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID2D, block_rows=tx * ty, dims=[nx, ny, 1, tx, ty, 1],
                            device_factor=dev_factor)
        M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=dtc, opts=o)
        info = M.info()
        assert info.device_factored == dev_factor and info.lattice == 1
        din, dout = dev(b, dt), dev(np.zeros(n), dt)
        M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
        outs.append(host(ctx, dout))
        infos.append((info.nnzL, info.nnzU, info.nblocks, info.max_block_rows, info.levelsL, info.stream_bytes))
        if dev_factor:  # and the solver runs on it (permuted matrix built from the device-produced 1/diag)
            r, x, _ = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, dtc, tol=1e-6 if dtc == 0 else 1e-8, max_iter=3000, orth=1)
            it_dev = r.iters
            assert r.flag == 0
        else:
            r, x, _ = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, dtc, tol=1e-6 if dtc == 0 else 1e-8, max_iter=3000, orth=1)
            assert r.iters == it_dev
        M.close()
    assert infos[0] == infos[1]
    assert np.array_equal(outs[0], outs[1])


def test_device_factorisation_declines_other_matrices(bg, ctx):
    """a matrix that is not a 5-point grid in the given numbering goes through the host path, silently and correctly"""
    rp, ci, v, b = CASES["rand_2000"]
    n = len(rp) - 1
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID2D, block_rows=256, dims=[50, 40, 1, 16, 16, 1])
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1, opts=o)
    assert M.info().device_factored == 0
    M.close()


@pytest.mark.parametrize("name", ["pg2d_24x20", "th3d_8x7x6"])
@pytest.mark.parametrize("dtc", [0, 1])
def test_split_preconditioner_matches_golden(bg, ctx, name, dtc):
    """MyILUPPfloat::DevPrecond_{left,right,starting_value} on ILU++ factors (committed fixture)."""
    g = load_golden(name + "_split")
    mode = "f32f64" if dtc == 0 else "f64"
    dt = np_dt(dtc)
    n = len(g["rowptr"]) - 1
    M = bg.Precond.split(ctx, g["lp"], g["li"], g["lv"], g["up"], g["ui"], g["uv"], g["middle"], g["lscale"], g["rscale"],
                         g["permRow"], g["permCol"], dtype=dtc)
    din, dout = dev(g["b"], dt), dev(np.zeros(n), dt)
    for which, nm in ((bg.PC_LEFT, "left"), (bg.PC_RIGHT, "right"), (bg.PC_START, "start"), (bg.PC_RHS, "left")):
        M.apply(which, din.data_ptr(), dout.data_ptr())
        got = host(ctx, dout)
        want = g[f"{mode}_{nm}_b"]
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < (2e-6 if dtc == 0 else 1e-13), nm
    M.close()


def test_diag_and_none(bg, ctx):
    rp, ci, v, b = CASES["pg2d_61x47"]
    n = len(rp) - 1
    d = sp.csr_matrix((v, ci, rp), shape=(n, n)).diagonal()
    din, dout = dev(b, np.float64), dev(np.zeros(n), np.float64)
    M = bg.Precond.diag(ctx, rp, ci, v)
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    assert np.allclose(host(ctx, dout), b / d, rtol=1e-15)
    M.close()
    M = bg.Precond.none(ctx, n)
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    assert np.array_equal(host(ctx, dout), b)
    M.close()


# ------------------------------------------------------------------------------- GMRES
def run_gpu(bg, ctx, rp, ci, v, b, x0, M, dtc, **kw):
    dt = np_dt(dtc)
    n = len(rp) - 1
    A = bg.Matrix(ctx, rp, ci, v, dtype=dtc)
    m = kw.get("m", 32)
    S = bg.Solver(ctx, n, m, dtype=dtc)
    db, dx = dev(b, dt), dev(x0, dt)
    r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), bg.gmres_opts(history_cap=4096, **kw))
    x = host(ctx, dx).astype(np.float64)
    hist = S.history()
    S.close()
    A.close()
    return r, x, hist


@pytest.mark.parametrize("schedule", ["level", "block"])
@pytest.mark.parametrize("shape", [(96, 80), (400, 300)])
def test_mixed_precision_basis_reaches_the_fp64_answer(bg, ctx, schedule, shape):
    """Opt-in FP32 Krylov basis under FP64 vectors (bg_solver_create_ex): same stopping rule, true residual still below the
    tolerance (the restart residual is always recomputed in FP64), solution within 1e-7 of the all-FP64 run, iteration
    count within 10 %; the combinations it does not cover are refused."""
    from b200gmres import synth

    nx, ny = shape
    rp, ci, v, b = synth.pg2d(nx, ny, seed=613)
    n = len(rp) - 1
    A_h = sp.csr_matrix((v, ci, rp), shape=(n, n))
    o = None if schedule == "level" else bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID2D, block_rows=1024,
                                                         dims=[nx, ny, 1, 32, 32, 1])
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1, opts=o)
    A = bg.Matrix(ctx, rp, ci, v, dtype=1)
    out = {}
    for basis in (bg.F64, bg.F32):
        S = bg.Solver(ctx, n, 32, dtype=bg.F64, basis_dtype=basis)
        for orth in (bg.ORTH_CGS_ADAPTIVE, bg.ORTH_CGS2):
            db, dx = dev(b, np.float64), dev(np.zeros(n), np.float64)
            r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), bg.gmres_opts(m=32, tol=1e-8, max_iter=5000, orth=orth, reorth_eta=0.1))
            x = host(ctx, dx)
            out[(basis, orth)] = (r, x, np.linalg.norm(b - A_h @ x) / np.linalg.norm(b))
        if basis == bg.F32:
            db, dx = dev(b, np.float64), dev(np.zeros(n), np.float64)
            with pytest.raises(bg.BgError):  # strict MGS keeps the basis in the vector type
                S.solve(A, M, db.data_ptr(), dx.data_ptr(), bg.gmres_opts(m=32, tol=1e-8, max_iter=10, orth=bg.ORTH_MGS))
        S.close()
    for orth in (bg.ORTH_CGS_ADAPTIVE, bg.ORTH_CGS2):
        (r64, x64, t64), (r32, x32, t32) = out[(bg.F64, orth)], out[(bg.F32, orth)]
        assert r64.flag == 0 and r32.flag == 0
        assert t32 <= 2e-8, f"true relative residual {t32:.3e}"
        assert abs(r32.iters - r64.iters) <= max(3, 0.1 * r64.iters), (r32.iters, r64.iters)
        assert np.linalg.norm(x32 - x64) <= 1e-7 * np.linalg.norm(x64)
    with pytest.raises(bg.BgError):
        bg.Solver(ctx, n, 64, dtype=bg.F64, basis_dtype=bg.F32)   # the fused CGS step covers restarts up to 32
    with pytest.raises(bg.BgError):
        bg.Solver(ctx, n, 32, dtype=bg.F32, basis_dtype=bg.F64)
    A.close(); M.close()


@pytest.mark.parametrize("dtc", [0, 1])
@pytest.mark.parametrize("shape", [(96, 80), (12, 9), (400, 300)])
def test_fused_mgs_step_equals_launch_pairs(bg, ctx, dtc, shape):
    """Strict MGS (the reference's order, src/gmres.cu:2356-2362) as ONE cooperative launch (mgs_kernel: w in registers,
    packet all-to-all sums) against the same step as i+1 dot/axpy launch pairs (BG_FUSED_ORTH=0): same stopping
    iteration, same Hessenberg column by column up to summation order, same solution.  108 rows: four CTAs, the last one partial;
    120 000 rows: every SM owns a slice; restart 5 exercises the restart path through the fused kernel."""
    from b200gmres import synth

    rp, ci, v, b = synth.pg2d(shape[0], shape[1], seed=977)
    n = len(rp) - 1
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=dtc)
    tol = 1e-5 if dtc == 0 else 1e-9
    res = {}
    for m in (32, 5):
        for fused in ("1", "0"):
            os.environ["BG_FUSED_ORTH"] = fused
            try:
                A = bg.Matrix(ctx, rp, ci, v, dtype=dtc)
                S = bg.Solver(ctx, n, m, dtype=dtc)
                assert S.mgs_fused() == (fused == "1")
                db, dx = dev(b, np_dt(dtc)), dev(np.zeros(n), np_dt(dtc))
                r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), bg.gmres_opts(m=m, tol=tol, max_iter=3000, orth=bg.ORTH_MGS, history_cap=4096))
                res[(m, fused)] = (r, host(ctx, dx).astype(np.float64), S.history(), S.hessenberg()[0])
                S.close(); A.close()
            finally:
                os.environ.pop("BG_FUSED_ORTH", None)
        (r1, x1, h1, H1), (r0, x0, h0, H0) = res[(m, "1")], res[(m, "0")]
        assert (r1.flag, r1.iters) == (r0.flag, r0.iters), f"m={m}: fused {r1.flag}/{r1.iters}/{r1.resid:.3e} vs pairs {r0.flag}/{r0.iters}/{r0.resid:.3e}"
        assert r1.flag == 0
        # rounding differences of the sums are amplified along the Krylov sequence: late columns agree to ~1e-8
        assert np.allclose(h1, h0, rtol=2e-3 if dtc == 0 else 1e-6, atol=1e-14)
        assert np.linalg.norm(x1 - x0) <= (5e-4 if dtc == 0 else 1e-9) * np.linalg.norm(x0)
        if m == 32:  # (short restarts end on a residual whose leading digits are rounding noise: H of that cycle is not comparable)
            assert np.allclose(H1, H0, rtol=0, atol=(1e-3 if dtc == 0 else 1e-6) * np.abs(H0).max())
    M.close()


@pytest.mark.parametrize("name", ["pg2d_24x20", "th3d_8x7x6", "rand_300"])
@pytest.mark.parametrize("orth", [0, 1, 2])
def test_gmres_ilu0_f64_matches_golden(bg, ctx, name, orth):
    g = load_golden(name)
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    flag, it, res, tol = g["f64_meta"]
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1)
    r, x, hist = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, 1, tol=tol, max_iter=2000, orth=orth)
    assert r.flag == int(flag) and r.iters == int(it)         # same stopping iteration as the reference loop
    assert abs(r.resid - res) <= RESID_ATOL_F64
    assert np.linalg.norm(x - g["f64_x"]) <= X_RTOL_F64 * np.linalg.norm(g["f64_x"])
    assert len(hist) == r.iters_done and np.allclose(hist, g["f64_hist"], rtol=1e-6, atol=1e-14)
    # restart path (m = 5) from a non-zero guess, in-out semantics of max_iter / tol
    flag, it, res, tol = g["f64_m5_meta"]
    r, x, _ = run_gpu(bg, ctx, rp, ci, v, b, np.ones(n), M, 1, tol=tol, max_iter=2000, orth=orth, m=5)
    assert r.flag == int(flag) and abs(r.iters - int(it)) <= 1 and r.restarts >= 1
    assert np.linalg.norm(x - g["f64_m5_x"]) <= 1e-7 * np.linalg.norm(g["f64_m5_x"])
    M.close()


@pytest.mark.parametrize("name", ["pg2d_24x20", "th3d_8x7x6"])
def test_gmres_none_and_diag_match_golden(bg, ctx, name):
    g = load_golden(name)
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    for key, M in (("none", bg.Precond.none(ctx, n)), ("diag", bg.Precond.diag(ctx, rp, ci, v))):
        flag, it, res, tol = g[key + "_meta"]
        r, x, _ = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, 1, tol=tol, max_iter=400, orth=0)
        assert r.flag == int(flag)
        if flag == 0:
            assert abs(r.iters - int(it)) <= 1 and np.linalg.norm(x - g[key + "_x"]) <= 1e-6 * np.linalg.norm(g[key + "_x"])
        M.close()


@pytest.mark.parametrize("name", ["pg2d_24x20", "th3d_8x7x6"])
def test_gmres_ilu0_f32_reference_precision(bg, ctx, name):
    """FP32 end to end like the reference (REAL float): same answer to FP32 accuracy."""
    g = load_golden(name)
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    flag, it, res, tol = g["f32_meta"]
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=0)
    r, x, hist = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, 0, tol=tol, max_iter=2000, orth=0)
    assert r.flag == 0 and r.resid < tol
    # the residual history follows the reference's own FP32 run (the oracle's float instantiation is bit-identical with the
    # reference's compiled GMRES()) for as long as float round-off does not dominate: 1e-3 relative while the residual
    # is above 1e-3, 2e-2 while it is above 1e-4 (the FP32 and FP64 oracle histories agree to 2e-5 / 3e-3 there)
    h32 = g["f32_hist"]
    k = min(len(hist), len(h32))
    for floor, rtol in ((1e-3, 1e-3), (1e-4, 2e-2)):
        sel = g["f64_hist"][:k] > floor
        assert np.allclose(np.asarray(hist)[:k][sel], h32[:k][sel], rtol=rtol, atol=0), (floor, np.abs(np.asarray(hist)[:k][sel] / h32[:k][sel] - 1).max())
    # FP32 round-off can only delay the stopping iteration relative to exact arithmetic: the FP32
    # oracle (float accumulators, like cublasSdot) may stagnate just above tol until the next restart
    # while the device dots accumulate in FP64.  So: no earlier than the FP64 trajectory crosses
    # tol, no later than the FP32 oracle (each with a small slack).
    first64 = int(np.argmax(g["f64_hist"] < tol)) + 1
    slack = max(3, int(0.15 * it))
    assert first64 - slack <= r.iters <= int(it) + slack
    assert np.linalg.norm(x - g["f64_x"]) <= 5e-4 * np.linalg.norm(g["f64_x"])
    M.close()


@pytest.mark.parametrize("name", ["pg2d_24x20", "th3d_8x7x6"])
@pytest.mark.parametrize("dtc", [0, 1])
def test_gmresilu_split_matches_golden(bg, ctx, name, dtc):
    """GMRESilu_GPU algebra (split preconditioning, warm start) on the reference's ILU++ factors."""
    g = load_golden(name + "_split")
    mode = "f32f64" if dtc == 0 else "f64"
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    M = bg.Precond.split(ctx, g["lp"], g["li"], g["lv"], g["up"], g["ui"], g["uv"], g["middle"], g["lscale"], g["rscale"],
                         g["permRow"], g["permCol"], dtype=dtc)
    flag, it, res, tol = g[f"{mode}_meta"]
    r, x, _ = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, dtc, tol=tol, max_iter=10000, orth=0)
    assert r.flag == 0
    if dtc == 1:
        assert r.iters == int(it) and abs(r.resid - res) <= RESID_ATOL_F64
        assert np.linalg.norm(x - g["f64_x"]) <= X_RTOL_F64 * np.linalg.norm(g["f64_x"])
    else:
        assert np.linalg.norm(x - g["f64_x"]) <= 2e-4 * np.linalg.norm(g["f64_x"])
    # warm start (transient pattern): x0 close to the solution, few iterations
    flag, it, res, tol = g[f"{mode}_warm_meta"]
    r, x, _ = run_gpu(bg, ctx, rp, ci, v, b, g[f"{mode}_warm_x0"], M, dtc, tol=tol, max_iter=10000, orth=0)
    assert r.flag == 0
    if dtc == 1:
        assert r.iters == int(it)
        assert np.linalg.norm(x - g["f64_warm_x"]) <= X_RTOL_F64 * np.linalg.norm(g["f64_warm_x"])
    M.close()


@pytest.mark.parametrize("name", ["pg2d_61x47", "th3d_13x11x9", "rand_2000"])
def test_gmres_matches_oracle_on_fresh_inputs(bg, ctx, oracle, name):
    rp, ci, v, b = CASES[name]
    n = len(rp) - 1
    f, pc = oracle_ilu0_pc(oracle, bg, rp, ci, v, "f64")
    flag, xo, it, res, hist = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 5000, 1e-8, pc, hist_cap=5000)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1)
    for orth in (0, 1):
        r, x, h = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, 1, tol=1e-8, max_iter=5000, orth=orth)
        assert (r.flag, r.iters) == (flag, it)
        assert abs(r.resid - res) <= RESID_ATOL_F64
        assert np.linalg.norm(x - xo) <= X_RTOL_F64 * np.linalg.norm(xo)
    M.close()


@pytest.mark.parametrize("name", ["pg2d_61x47", "th3d_13x11x9", "rand_2000"])
@pytest.mark.parametrize("level", [1, 2])
def test_gmres_iluk_matches_oracle(bg, ctx, oracle, name, level):
    """ILU(k) factors (bg_precond_create_iluk: ITSOL's level-of-fill rule, pinned on the host against the code the
    reference vendors) through the general reference-order triangular solves: same iteration count as the oracle
    given the same factors, fewer iterations than ILU(0)."""
    rp, ci, v, b = CASES[name]
    n = len(rp) - 1
    f = bg.host_iluk(rp, ci, v, level)
    pc = oracle.Precond(oracle.PC_ILU0, n, "f64", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
    flag, xo, it, res, hist = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 5000, 1e-8, pc, hist_cap=5000)
    _, pc0 = oracle_ilu0_pc(oracle, bg, rp, ci, v, "f64")
    it0 = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 5000, 1e-8, pc0)[2]
    assert it < it0
    M = bg.Precond.iluk(ctx, rp, ci, v, level, dtype=1)
    r, x, h = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, 1, tol=1e-8, max_iter=5000, orth=0)
    assert (r.flag, r.iters) == (flag, it)
    assert abs(r.resid - res) <= RESID_ATOL_F64
    assert np.linalg.norm(x - xo) <= X_RTOL_F64 * np.linalg.norm(xo)
    M.close()


def test_block_jacobi_gmres_converges_and_reports_drift(bg, ctx, oracle):
    rp, ci, v, b = CASES["pg2d_61x47"]
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    xs = sp.linalg.spsolve(A.tocsc(), b)
    M0 = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1)
    r0, x0, _ = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M0, 1, tol=1e-8, max_iter=5000)
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID2D, block_rows=256, dims=[61, 47, 1, 16, 16, 1])
    M1 = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1, opts=o)
    r1, x1, _ = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M1, 1, tol=1e-8, max_iter=5000)
    assert r0.flag == 0 and r1.flag == 0
    assert r0.iters <= r1.iters <= 2 * r0.iters      # block-Jacobi drift is bounded and reported
    for x in (x0, x1):
        assert np.linalg.norm(A @ x - b) <= 1e-6 * np.linalg.norm(b)
        assert np.linalg.norm(x - xs) <= 1e-5 * np.linalg.norm(xs)
    # the oracle with the same block-Jacobi factors stops at the same iteration
    nb, blk = bg.host_block_partition(bg.PART_GRID2D, rp, ci, 0, [61, 47, 1, 16, 16, 1])
    _, pc = oracle_ilu0_pc(oracle, bg, rp, ci, v, "f64", blk)
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 5000, 1e-8, pc)
    assert r1.iters == it and np.linalg.norm(x1 - xo) <= X_RTOL_F64 * np.linalg.norm(xo)
    M0.close()
    M1.close()


def test_gmres_edge_cases(bg, ctx):
    g = load_golden("pg2d_24x20")
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1)
    # exact initial guess: returns at the first check with 0 iterations ("<=", gmres.cu:2312-2314)
    r, x, _ = run_gpu(bg, ctx, rp, ci, v, b, g["f64_x"], M, 1, tol=1e-6, max_iter=100)
    assert (r.flag, r.iters, r.iters_done) == (0, 0, 0)
    # zero right-hand side: normb falls back to 1, x stays 0
    r, x, _ = run_gpu(bg, ctx, rp, ci, v, np.zeros(n), np.zeros(n), M, 1, tol=1e-8, max_iter=100)
    assert r.flag == 0 and r.iters == 0 and not x.any()
    # iteration cap: flag 1, achieved residual reported through tol
    r, x, _ = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, 1, tol=1e-12, max_iter=7)
    assert r.flag == 1 and r.iters_done == 7 and r.resid > 1e-12
    # host-buffer entry point (GMRES_dev_PG): H2D, solve, D2H
    A = bg.Matrix(ctx, rp, ci, v, dtype=1)
    S = bg.Solver(ctx, n, 32, dtype=1)
    xh = np.zeros(n)
    r = S.solve(A, M, np.ascontiguousarray(b), xh, bg.gmres_opts(tol=g["f64_meta"][3], max_iter=2000, orth=0), host=True)
    assert r.iters == int(g["f64_meta"][1]) and np.linalg.norm(xh - g["f64_x"]) <= X_RTOL_F64 * np.linalg.norm(g["f64_x"])
    # workspace reuse across solves (GMRES_GPU_tran pattern): second solve identical
    xh2 = np.zeros(n)
    r2 = S.solve(A, M, np.ascontiguousarray(b), xh2, bg.gmres_opts(tol=g["f64_meta"][3], max_iter=2000, orth=0), host=True)
    assert r2.iters == r.iters and np.array_equal(xh, xh2)
    S.close()
    A.close()
    M.close()


def test_arnoldi_relation_at_scale(bg, ctx):
    """Size-independent property at a larger size: after one full cycle the columns of V are
    orthonormal and the Givens residual estimate equals the true preconditioned residual."""
    from b200gmres import synth

    rp, ci, v, b = synth.pg2d(300, 280, seed=211)
    n = len(rp) - 1
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID2D, block_rows=1024, dims=[300, 280, 1, 32, 32, 1])
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=1, opts=o)
    r, x, hist = run_gpu(bg, ctx, rp, ci, v, b, np.zeros(n), M, 1, tol=1e-8, max_iter=3000)
    assert r.flag == 0
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    # final estimate vs the true preconditioned relative residual ||M^-1(b-Ax)|| / ||M^-1 b||
    dt = np.float64
    dr, dz, db, dzb = dev(b - A @ x, dt), dev(np.zeros(n), dt), dev(b, dt), dev(np.zeros(n), dt)
    M.apply(bg.PC_FULL, dr.data_ptr(), dz.data_ptr())
    M.apply(bg.PC_FULL, db.data_ptr(), dzb.data_ptr())
    true = np.linalg.norm(host(ctx, dz)) / np.linalg.norm(host(ctx, dzb))
    assert true < 2e-8 and abs(true - r.resid) < 1e-9
    assert np.all(np.diff(hist[:32]) <= 1e-15)  # residual estimates are monotone inside a cycle
    M.close()
