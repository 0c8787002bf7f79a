"""The CPU oracle against the committed golden vectors and against the reference's own CPU
translation units (oracle/_ref, built only where /root/reference exists)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden

CASES = ["pg2d_24x20", "th3d_8x7x6", "rand_300"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["f32", "f64"])
def test_ilu0_and_gmres_match_golden(oracle, name, mode):
    g = load_golden(name)
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    f = oracle.ilu0(rp, ci, v, mode)
    for k in ("lp", "li", "lv", "up", "ui", "uv", "level", "levelPtr", "nodes"):
        assert np.array_equal(f[k], g[f"{mode}_{k}"]), k  # bit-exact, values included
    assert np.array_equal(oracle.get_level(1, f["lp"], f["li"], f["lv"], mode), g[f"{mode}_levL"])
    assert np.array_equal(oracle.get_level(0, f["up"], f["ui"], f["uv"], mode), g[f"{mode}_levU"])
    pc = oracle.Precond(oracle.PC_ILU0, n, mode, lp=f["lp"], li=f["li"], lv=f["lv"], up=f["up"], ui=f["ui"], uv=f["uv"])
    flag, it, res, tol = g[f"{mode}_meta"]
    fl, x, its, rs, hist = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 2000, tol, pc, hist_cap=2000)
    assert (fl, its) == (int(flag), int(it))
    assert np.array_equal(x, g[f"{mode}_x"]) and np.array_equal(hist, g[f"{mode}_hist"])
    assert np.array_equal(pc.apply(0, b), g[f"{mode}_Minv_b"])


@pytest.mark.parametrize("name", CASES)
def test_ilu0_is_textbook_ilu0(oracle, name):
    """(L U - A) vanishes on the pattern of A: pins the restated leftILU against the definition."""
    g = load_golden(name)
    rp, ci, v = g["rowptr"], g["col"], g["val"]
    n = len(rp) - 1
    f = oracle.ilu0(rp, ci, v, "f64")
    L = sp.csr_matrix((f["lv"], f["li"], f["lp"]), shape=(n, n))
    U = sp.csr_matrix((f["uv"], f["ui"], f["up"]), shape=(n, n))
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    E = (L @ U - A).multiply(A != 0)
    assert abs(E).max() < 1e-12 * abs(A).max()
    # reference layout: L diagonal last and unit, U diagonal first
    assert np.array_equal(f["li"][f["lp"][1:] - 1], np.arange(n)) and np.all(f["lv"][f["lp"][1:] - 1] == 1.0)
    assert np.array_equal(f["ui"][f["up"][:-1]], np.arange(n))


@pytest.mark.parametrize("name", CASES)
def test_gmres_solves_the_system(oracle, name):
    g = load_golden(name)
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    xs = sp.linalg.spsolve(A.tocsc(), b)
    assert np.linalg.norm(g["f64_x"] - xs) <= 1e-6 * np.linalg.norm(xs)
    assert np.linalg.norm(g["f64_m5_x"] - xs) <= 1e-6 * np.linalg.norm(xs)
    assert np.linalg.norm(g["none_x"] - xs) <= 1e-4 * np.linalg.norm(xs) or g["none_meta"][0] == 1


@pytest.mark.parametrize("name", ["pg2d_24x20", "th3d_8x7x6"])
def test_split_preconditioner_golden(oracle, name):
    g = load_golden(name + "_split")
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    keys = ("lp", "li", "lv", "up", "ui", "uv", "middle", "lscale", "rscale", "permRow", "permCol")
    for mode in ("f32f64", "f64"):
        pc = oracle.Precond(oracle.PC_SPLIT, n, mode, **{k: g[k] for k in keys})
        flag, it, res, tol = g[f"{mode}_meta"]
        fl, x, its, rs, hist = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, tol, pc, hist_cap=1000)
        assert (fl, its) == (int(flag), int(it)) and np.array_equal(x, g[f"{mode}_x"])
        for which, nm in ((1, "left"), (2, "right"), (3, "start")):
            assert np.array_equal(pc.apply(which, b), g[f"{mode}_{nm}_b"])
        # right o start = identity: the warm-start transform inverts M_R^-1 (preconditioner.cu:1561-1589)
        y = pc.apply(3, g[f"{mode}_x"])
        back = pc.apply(2, y)
        assert np.linalg.norm(back - g[f"{mode}_x"]) <= (1e-4 if mode == "f32f64" else 1e-10) * np.linalg.norm(g[f"{mode}_x"])
    # the f64 oracle loop and ILU++'s own GMRES agree on the solution of the same system
    assert np.linalg.norm(g["ilupp_x"] - g["f64_x"]) <= 1e-6 * np.linalg.norm(g["f64_x"])


# ---------------------------------------------------------------- against the reference's own code
def _need_ref(oracle):
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return r


@pytest.mark.parametrize("name", CASES)
def test_spmv_and_lusolve_bit_exact_with_reference_sources(oracle, name):
    import ctypes as C

    ref = _need_ref(oracle)
    g = load_golden(name)
    rp, ci = g["rowptr"].astype(np.int32), g["col"].astype(np.int32)
    v, b = g["val"].astype(np.float32), g["b"].astype(np.float32)
    n = len(rp) - 1
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    y = np.zeros(n, np.float32)
    ref.ref_computeSpMV(p(y), p(v), p(rp), p(ci), p(b), n)  # src/SpMV_compute.cpp:19-37
    assert np.array_equal(y, oracle.spmv(rp, ci, v, b, "f32"))
    f = {k: g["f32_" + k] for k in ("lp", "li", "lv", "up", "ui", "uv")}
    x = np.zeros(n, np.float32)
    ref.ref_LUSolve_ignoreZero(p(x), p(f["lv"]), p(f["lp"]), p(f["li"]), p(f["uv"]), p(f["up"]), p(f["ui"]), p(b), n)
    pc = oracle.Precond(oracle.PC_ILU0, n, "f32", **f)
    assert np.array_equal(x, pc.apply(0, b))  # MyILU0::HostPrecond == LUSolve_ignoreZero


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("npart", [2, 4, 7])
def test_partition_bit_exact_with_reference_sources(oracle, name, npart):
    import ctypes as C

    ref = _need_ref(oracle)
    g = load_golden(name)
    n = len(g["rowptr"]) - 1
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]), shape=(n, n)).tocsc()
    A.sort_indices()
    Ap, Ai = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    ps, npt, pinv, q = (np.zeros(npart + 1, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64))
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    ref.ref_partition(C.c_long(n), p(Ap), p(Ai), npart, n, p(ps), p(npt), p(pinv), p(q))  # src/partition.cpp:22-99
    got = oracle.partition(Ap, Ai, npart, n)
    for a, b in zip((ps, npt, pinv, q), got):
        assert np.array_equal(a, b)
    if npart in (2, 4):
        assert np.array_equal(ps, g[f"part{npart}_size"]) and np.array_equal(q, g[f"part{npart}_q"])


def test_ilupp_factors_reproduce_golden(oracle):
    """ILU++ run here equals the committed fixture (ties the fixture to the vendored library)."""
    if oracle.ilupp() is None:
        pytest.skip("oracle/_ref/libilupp.so not built (needs /root/reference)")
    g = load_golden("pg2d_24x20_split")
    n = len(g["rowptr"]) - 1
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]), shape=(n, n)).tocsc()
    A.sort_indices()
    fac = oracle.ilupp_factor(n, A.indptr, A.indices, A.data)
    for k in ("lp", "li", "lv", "up", "ui", "uv", "middle", "permRow", "permCol", "lscale", "rscale"):
        assert np.array_equal(fac[k], g[k]), k
    oracle.ilupp_free(fac)
