"""GPU parity at the BASELINE.json configuration sizes (B200): the CUDA path, called through the C ABI,
against the CPU oracle (oracle/oracle.cpp = the reference's host GMRES / GMRESilu, src/gmres.cu:1904-2066,
2069-2252) on the same seeded inputs.

    C1  pg2d 174^2  (ibmpg1-sized, configs[0])          level schedule, strict MGS and fused CGS
    C2  pg2d 1304^2 (ibmpg6-sized, configs[1])          level schedule (reference operator) AND block-Jacobi
                                                         (the oracle is given the same block_of_row)
    C2' pg2d 300^2 / 1304^2 + ILU++ factors             GMRESilu_GPU split path (a1/a6/a7/a20)
    C3  th3d 64^3  (7-point thermal, configs[2] shape)  level schedule (sync-free kernel) and 3-D tiles

Bars (north_star): iteration count identical, |resid - resid_oracle| <= 1e-10 on the relative residual,
||x - x_oracle|| <= 1e-8 ||x_oracle||, level sets bit-exact.  The oracle runs take 0.2 s (C1) to ~20 s (C2)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

X_RTOL = 1e-8
RESID_ATOL = 1e-10


def dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    torch.cuda.synchronize()
    return t


def gpu_solve(bg, ctx, rp, ci, v, b, M, **kw):
    n = len(rp) - 1
    A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
    S = bg.Solver(ctx, n, kw.get("m", 32), dtype=bg.F64)
    db, dx = dev(b), dev(np.zeros(n))
    r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), bg.gmres_opts(**kw))
    ctx.sync()
    x = dx.cpu().numpy()
    S.close()
    A.close()
    return r, x


def ilu0_oracle(oracle, bg, rp, ci, v, blk=None):
    f = bg.host_ilu0(rp, ci, v, blk)  # bit-exact with the reference's leftILU (tests/test_host_logic.py)
    n = len(rp) - 1
    return f, oracle.Precond(oracle.PC_ILU0, n, "f64", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})


def check(r, x, flag, it, res, xo, what):
    assert (r.flag, r.iters) == (flag, it), f"{what}: flag/iters {r.flag}/{r.iters} vs oracle {flag}/{it}"
    assert abs(r.resid - res) <= RESID_ATOL, f"{what}: resid {r.resid} vs {res}"
    err = np.linalg.norm(x - xo) / np.linalg.norm(xo)
    assert err <= X_RTOL, f"{what}: |x - x_oracle|/|x| = {err:.3e}"
    return err


@pytest.fixture(scope="module")
def c2(bg, oracle):
    from b200gmres import synth

    rp, ci, v, b = synth.pg2d(1304, 1304, seed=synth.SEED_BASE + 2)  # bench.py's configs[1] input
    return rp, ci, v, b


def test_c1_pg2d_174_level_matches_oracle(bg, ctx, oracle):
    from b200gmres import synth

    rp, ci, v, b = synth.pg2d(174, 174, seed=synth.SEED_BASE + 1)
    n = len(rp) - 1
    f, pc = ilu0_oracle(oracle, bg, rp, ci, v)
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-8, pc)
    for lattice in (1, 0):  # wavefront strips / general sync-free slices: the same operator
        M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=bg.precond_opts(lattice=lattice))
        assert M.info().lattice == lattice
        assert np.array_equal(M.levels(True), oracle.get_level(1, f["lp"], f["li"], f["lv"], "f64"))
        assert np.array_equal(M.levels(False), oracle.get_level(0, f["up"], f["ui"], f["uv"], "f64"))
        for orth in (bg.ORTH_MGS, bg.ORTH_CGS_ADAPTIVE):
            r, x = gpu_solve(bg, ctx, rp, ci, v, b, M, tol=1e-8, max_iter=10000, orth=orth)
            check(r, x, flag, it, res, xo, f"C1 lattice={lattice} orth={orth}")
        M.close()


def test_c2_pg2d_1304_level_matches_oracle(bg, ctx, oracle, c2):
    rp, ci, v, b = c2
    n = len(rp) - 1
    f, pc = ilu0_oracle(oracle, bg, rp, ci, v)
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-8, pc)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64)
    info = M.info()
    assert info.lattice == 1 and info.levelsL == 2 * 1304 - 1
    assert np.array_equal(M.levels(True), oracle.get_level(1, f["lp"], f["li"], f["lv"], "f64"))
    assert np.array_equal(M.levels(False), oracle.get_level(0, f["up"], f["ui"], f["uv"], "f64"))
    # one application of the reference operator at full size
    din, dout = dev(b), dev(np.zeros(n))
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    ctx.sync()
    want = pc.apply(0, b)
    assert np.linalg.norm(dout.cpu().numpy() - want) <= 1e-13 * np.linalg.norm(want)
    for orth in (bg.ORTH_MGS, bg.ORTH_CGS_ADAPTIVE):
        r, x = gpu_solve(bg, ctx, rp, ci, v, b, M, tol=1e-8, max_iter=10000, orth=orth)
        check(r, x, flag, it, res, xo, f"C2 level orth={orth}")
    M.close()
    # the general sync-free kernel on the same factors (what an ILUT / ILU++ pattern of this size would run)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=bg.precond_opts(lattice=0))
    dout.zero_()
    M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
    ctx.sync()
    assert np.linalg.norm(dout.cpu().numpy() - want) <= 1e-14 * np.linalg.norm(want)
    r, x = gpu_solve(bg, ctx, rp, ci, v, b, M, tol=1e-8, max_iter=10000, orth=bg.ORTH_CGS_ADAPTIVE)
    check(r, x, flag, it, res, xo, "C2 sync-free")
    M.close()


def test_c2_pg2d_1304_block_matches_oracle_with_same_blocks(bg, ctx, oracle, c2):
    rp, ci, v, b = c2
    n = len(rp) - 1
    dims = [1304, 1304, 1, 32, 32, 1]
    nb, blk = bg.host_block_partition(bg.PART_GRID2D, rp, ci, 0, dims)
    _, pc = ilu0_oracle(oracle, bg, rp, ci, v, blk)
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-8, pc)
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID2D, block_rows=1024, dims=dims)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o)
    assert M.info().nblocks == nb
    for orth in (bg.ORTH_MGS, bg.ORTH_CGS_ADAPTIVE):
        r, x = gpu_solve(bg, ctx, rp, ci, v, b, M, tol=1e-8, max_iter=10000, orth=orth)
        check(r, x, flag, it, res, xo, f"C2 block orth={orth}")
    M.close()


@pytest.mark.parametrize("nx", [300, 1304])
def test_c2_split_ilupp_factors_match_oracle(bg, ctx, oracle, nx, c2):
    """GMRESilu_GPU on the factors the reference's own ILU++ produces (threshold 1.2,
    src/mna_solve_gpu_gmres.cpp:316-344), at 300^2 and at the ibmpg6 size."""
    if oracle.ilupp() is None:
        pytest.skip("oracle/_ref/libilupp.so not present")
    from b200gmres import synth

    rp, ci, v, b = c2 if nx == 1304 else synth.pg2d(nx, nx, seed=synth.SEED_BASE + 5)
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc()
    A.sort_indices()
    fac = oracle.ilupp_factor(n, A.indptr, A.indices, A.data, threshold=1.2, mem_factor=1.0)
    keys = ("lp", "li", "lv", "up", "ui", "uv", "middle", "lscale", "rscale", "permRow", "permCol")
    pc = oracle.Precond(oracle.PC_SPLIT, n, "f64", **{k: fac[k] for k in keys})
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-8, pc)
    M = bg.Precond.split(ctx, *[fac[k] for k in keys], dtype=bg.F64)
    for which, nm in ((bg.PC_LEFT, 1), (bg.PC_RIGHT, 2), (bg.PC_START, 3)):
        din, dout = dev(b), dev(np.zeros(n))
        M.apply(which, din.data_ptr(), dout.data_ptr())
        ctx.sync()
        want = pc.apply(nm, b)
        assert np.linalg.norm(dout.cpu().numpy() - want) <= 1e-13 * np.linalg.norm(want), nm
    r, x = gpu_solve(bg, ctx, rp, ci, v, b, M, tol=1e-8, max_iter=10000, orth=bg.ORTH_MGS)
    check(r, x, flag, it, res, xo, f"split {nx}^2")
    M.close()
    oracle.ilupp_free(fac)


def test_netlist_mna_system_through_the_power_grid_path(bg, ctx, oracle):
    """The reference's power-grid flow from the netlist on: parser -> full MNA system (voltage-source rows with a zero
    diagonal, src/mna.cpp:204-218) -> ILU++ factors with pivoting (threshold 1.2) -> setPrecondPG / GMRESilu_GPU.
    bg_netlist_* (bit-exact with the reference's parser, tests/test_netlist.py) feeds bg_precond_create_split; the
    device solve equals the oracle's GMRESilu on the same factors."""
    if oracle.ilupp() is None:
        pytest.skip("oracle/_ref/libilupp.so not present")
    from b200gmres import synth

    N = bg.Netlist(synth.pg2d_netlist(40, 36, seed=synth.SEED_BASE + 31, pad_pitch=8), text=True)
    n = N.info.size
    rp, ci, v = N.G
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    assert int((A.diagonal() == 0).sum()) == N.info.n_vsources > 0
    b = sp.csr_matrix((N.B[2], N.B[1], N.B[0]), shape=(n, N.info.n_sources)) @ N.u(0.0)
    Ac = A.tocsc()
    Ac.sort_indices()
    fac = oracle.ilupp_factor(n, Ac.indptr, Ac.indices, Ac.data, threshold=1.2, mem_factor=1.0)
    keys = ("lp", "li", "lv", "up", "ui", "uv", "middle", "lscale", "rscale", "permRow", "permCol")
    pc = oracle.Precond(oracle.PC_SPLIT, n, "f64", **{k: fac[k] for k in keys})
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-8, pc)
    assert flag == 0
    M = bg.Precond.split(ctx, *[fac[k] for k in keys], dtype=bg.F64)
    r, x = gpu_solve(bg, ctx, rp, ci, v, b, M, tol=1e-8, max_iter=10000, orth=bg.ORTH_MGS)
    check(r, x, flag, it, res, xo, "netlist MNA + ILU++")
    M.close()
    oracle.ilupp_free(fac)


def test_c3_th3d_64_level_and_tiles_match_oracle(bg, ctx, oracle):
    from b200gmres import synth

    rp, ci, v, b = synth.th3d(64, 64, 64, seed=synth.SEED_BASE + 3)
    n = len(rp) - 1
    f, pc = ilu0_oracle(oracle, bg, rp, ci, v)
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-8, pc)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64)
    assert np.array_equal(M.levels(True), oracle.get_level(1, f["lp"], f["li"], f["lv"], "f64"))
    for orth in (bg.ORTH_MGS, bg.ORTH_CGS_ADAPTIVE):
        r, x = gpu_solve(bg, ctx, rp, ci, v, b, M, tol=1e-8, max_iter=10000, orth=orth)
        check(r, x, flag, it, res, xo, f"C3 level orth={orth}")
    M.close()
    dims = [64, 64, 64, 8, 8, 8]
    nb, blk = bg.host_block_partition(bg.PART_GRID3D, rp, ci, 0, dims)
    _, pcb = ilu0_oracle(oracle, bg, rp, ci, v, blk)
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-8, pcb)
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID3D, block_rows=512, dims=dims)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o)
    r, x = gpu_solve(bg, ctx, rp, ci, v, b, M, tol=1e-8, max_iter=10000, orth=bg.ORTH_CGS_ADAPTIVE)
    check(r, x, flag, it, res, xo, "C3 3-D tiles")
    M.close()


def test_headline_10m_size_independent_properties(bg, ctx):
    """The headline workload itself (pg2d 3162^2 = 10 M rows, bench.py's input), where a CPU solve takes a minute: the
    CUDA path is checked through properties that do not need one.
      * SpMV against scipy on the same CSR (every row, 1e-13);
      * block-Jacobi preconditioner (factored AND packed on the device): linearity M(a u + b v) = a M u + b M v, and the
        defining identity L U (M r) = r with the HOST factors of the same tiles (bit-exact twin of the device ones);
      * the global (reference-order) ILU(0) through the wavefront-strip kernel: L U (M r) = r with the host factors;
      * the solve: true relative residual below the tolerance, the pinned iteration count of this input (108 / 84),
        and the solutions of the two operators agree to the accuracy the tolerance allows."""
    from b200gmres import synth

    nx = 3162
    rp, ci, v, b = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 2)
    n = len(rp) - 1
    A_h = sp.csr_matrix((v, ci, rp), shape=(n, n))
    rng = np.random.default_rng(1)
    u, w = rng.standard_normal(n), rng.standard_normal(n)
    A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
    du, dy = dev(u), dev(np.zeros(n))
    A.spmv(du.data_ptr(), dy.data_ptr())
    ctx.sync()
    want = A_h @ u
    assert np.abs(dy.cpu().numpy() - want).max() <= 1e-13 * np.abs(want).max()

    def lu_product(f, z):
        L = sp.csr_matrix((f["lv"], f["li"], f["lp"]), shape=(n, n))
        U = sp.csr_matrix((f["uv"], f["ui"], f["up"]), shape=(n, n))
        return L @ (U @ z)

    dims = [nx, nx, 1, 32, 32, 1]
    nb, blk = bg.host_block_partition(bg.PART_GRID2D, rp, ci, 0, dims)
    xs = {}
    for name, opts, f, iters in (
            ("block", bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_method=bg.PART_GRID2D, block_rows=1024, dims=dims), bg.host_ilu0(rp, ci, v, blk), 108),
            ("level", None, bg.host_ilu0(rp, ci, v), 84)):
        M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=opts)
        info = M.info()
        assert info.lattice == 1 and (name == "level" or (info.device_factored == 1 and info.nblocks == nb))
        outs = []
        for vec in (u, w, 0.75 * u - 2.5 * w):
            din, dout = dev(vec), dev(np.zeros(n))
            M.apply(bg.PC_FULL, din.data_ptr(), dout.data_ptr())
            ctx.sync()
            outs.append(dout.cpu().numpy())
        zu, zw, zc = outs
        assert np.linalg.norm(zc - (0.75 * zu - 2.5 * zw)) <= 1e-13 * np.linalg.norm(zc), name          # linearity
        assert np.linalg.norm(lu_product(f, zu) - u) <= 1e-12 * np.linalg.norm(u), name                 # L U z = r
        S = bg.Solver(ctx, n, 32, dtype=bg.F64)
        db, dx = dev(b), dev(np.zeros(n))
        r = S.solve(A, M, db.data_ptr(), dx.data_ptr(), bg.gmres_opts(m=32, tol=1e-8, max_iter=10000, orth=bg.ORTH_CGS_ADAPTIVE, reorth_eta=0.1))
        ctx.sync()
        x = dx.cpu().numpy()
        S.close(); M.close()
        assert r.flag == 0 and r.iters == iters, (name, r.iters)
        assert np.linalg.norm(b - A_h @ x) <= 1e-8 * np.linalg.norm(b), name
        xs[name] = x
    A.close()
    assert np.linalg.norm(xs["block"] - xs["level"]) <= 1e-6 * np.linalg.norm(xs["level"])
