"""Product host code (no GPU needed): ILU(0) pattern/values, level sets, partition vectors are
bit-exact with the oracle and the golden fixtures; the C ABI library loads and exports every
symbol include/b200gmres.h declares; compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden

CASES = ["pg2d_24x20", "th3d_8x7x6", "rand_300"]


def test_library_exports_every_declared_symbol(bg):
    names = re.findall(r"BG_API\s+[\w\s\*]+?\b(bg_\w+)\s*\(", open(bg.HEADER).read())
    assert len(names) > 40
    L = C.CDLL(bg.LIB_PATH)
    missing = [n for n in set(names) if not hasattr(L, n)]
    assert not missing, missing


def test_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: include/b200gmres.h must compile as C99 (no C++ types in the signatures)"""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "b200gmres.h"\nint main(void) { bg_gmres_opts o; bg_netlist_info i; (void)o; (void)i; return BG_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), "-fsyntax-only", str(src)])


def test_c_example_builds_and_fails_loudly_without_a_gpu(bg, tmp_path):
    """examples/solve_csr.c: plain C99 against the C ABI; without a device it must stop at bg_context_create with the
    library's message (exit status 3) -- never compute on the CPU; with a device it solves (status 0)"""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "solve_csr")
    libdir = os.path.join(root, "gpu-gmres_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "solve_csr.c"),
                           "-L", libdir, "-lb200gmres", "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    out = subprocess.run([exe, "40"], capture_output=True, text=True)
    assert out.returncode in (0, 3), out.stdout + out.stderr
    if out.returncode == 3:
        assert "no CPU fallback" in out.stderr
    else:
        assert "flag 0" in out.stdout


def test_no_cpu_fallback(bg):
    if bg.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(bg.BgError) as e:
        bg.Context(0)
    assert e.value.code == bg.ERR_NO_GPU


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_ilu0_bit_exact(bg, name, dt):
    g = load_golden(name)
    mode = "f32" if dt == np.float32 else "f64"
    f = bg.host_ilu0(g["rowptr"], g["col"], g["val"].astype(dt))
    for k in ("lp", "li", "up", "ui"):
        assert np.array_equal(f[k], g[f"{mode}_{k}"]), k  # pattern: integer artefact
    assert np.array_equal(f["lv"], g[f"{mode}_lv"]) and np.array_equal(f["uv"], g[f"{mode}_uv"])  # same op order
    assert f["zero_pivots"] == 0
    assert np.array_equal(bg.host_trsv_levels(True, f["lp"], f["li"], f["lv"]), g[f"{mode}_levL"])
    assert np.array_equal(bg.host_trsv_levels(False, f["up"], f["ui"], f["uv"]), g[f"{mode}_levU"])


@pytest.mark.parametrize("name", CASES)
def test_factor_levels_bit_exact(bg, name):
    g = load_golden(name)
    lev = bg.host_factor_levels(g["rowptr"], g["col"], g["val"])
    assert np.array_equal(lev, g["f32_level"])
    nl, lp, nodes = bg.host_bucket_levels(lev)
    assert np.array_equal(lp[: nl + 1], g["f32_levelPtr"]) and np.array_equal(nodes, g["f32_nodes"])


def test_structural_zero_rule(bg, oracle):
    """|v| < 1e-9 entries neither create levels nor survive the L/U split."""
    g = load_golden("pg2d_24x20")
    v = g["val"].copy()
    rp, ci = g["rowptr"], g["col"]
    off = np.flatnonzero(ci != np.repeat(np.arange(len(rp) - 1), np.diff(rp)))
    v[off[::7]] = 5e-10
    fo = oracle.ilu0(rp, ci, v, "f32")
    fp = bg.host_ilu0(rp, ci, v.astype(np.float32))
    for k in ("lp", "li", "lv", "up", "ui", "uv"):
        assert np.array_equal(fo[k], fp[k]), k
    assert len(fp["lv"]) < len(g["f32_lv"])
    assert np.array_equal(oracle.generate_level(rp, ci, v, "f32")[0], bg.host_factor_levels(rp, ci, v))


def test_unsorted_and_empty_rows(bg, oracle):
    rp = np.array([0, 2, 2, 5, 6], np.int32)  # row 1 empty, row 2 unsorted
    ci = np.array([0, 2, 3, 0, 2, 3], np.int32)
    v = np.array([4.0, 1.0, 1.0, 2.0, 5.0, 3.0])
    fo = oracle.ilu0(rp, ci, v, "f64")
    fp = bg.host_ilu0(rp, ci, v)
    for k in ("lp", "li", "lv", "up", "ui", "uv"):
        assert np.array_equal(fo[k], fp[k]), k
    assert fp["zero_pivots"] == 1  # the empty row has no pivot


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("npart", [2, 4, 5])
def test_partition_bit_exact(bg, oracle, name, npart):
    g = load_golden(name)
    n = len(g["rowptr"]) - 1
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]), shape=(n, n)).tocsc()
    A.sort_indices()
    got = bg.host_partition(A.indptr, A.indices, npart, n)
    want = oracle.partition(A.indptr, A.indices, npart, n)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    if npart in (2, 4):
        assert np.array_equal(got[1], g[f"part{npart}_node"]) and np.array_equal(got[2], g[f"part{npart}_pinv"])
    # extra unknowns beyond nNodes inherit a part (partition.cpp:56-59)
    got = bg.host_partition(A.indptr, A.indices, npart, n - 7)
    want = oracle.partition(A.indptr, A.indices, npart, n - 7)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_partition3_post(bg):
    g = load_golden("pg2d_24x20")
    n = len(g["rowptr"]) - 1
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]), shape=(n, n)).tocsc()
    part = (np.arange(n) % 24 >= 12).astype(np.int64)
    ps, npt, pinv, q = bg.host_partition3_post(A.indptr, A.indices, 2, part)
    cut = np.zeros(n, bool)
    C_ = A.tocoo()
    m = part[C_.row] != part[C_.col]
    cut[C_.row[m]] = cut[C_.col[m]] = True  # both endpoints (partition3.cpp:79-90)
    assert np.array_equal(npt == 2, cut) and ps.sum() == n and np.array_equal(q[pinv], np.arange(n))


def test_block_partitions(bg):
    g = load_golden("pg2d_24x20")
    rp, ci = g["rowptr"], g["col"]
    nb, blk = bg.host_block_partition(bg.PART_GRID2D, rp, ci, 0, [24, 20, 1, 8, 5, 1])
    assert nb == 12 and np.all(np.bincount(blk) == 40)
    nb, blk = bg.host_block_partition(bg.PART_METIS, rp, ci, 6)
    cnt = np.bincount(blk, minlength=6)
    assert nb == 6 and cnt.min() > 60 and cnt.max() < 100
    nb, blk = bg.host_block_partition(bg.PART_CONTIGUOUS, rp, ci, 7)
    assert np.all(np.diff(blk) >= 0) and blk.max() == 6
    # block-Jacobi ILU(0): no factor entry couples two blocks
    f = bg.host_ilu0(rp, ci, g["val"], blk)
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(f["lp"]))
    assert np.all(blk[rows] == blk[f["li"]])


def test_ilut_reduces_to_ilu_and_improves(bg, oracle):
    g = load_golden("pg2d_24x20")
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    f = bg.host_ilut(rp, ci, v, 0.0, n)  # no dropping, unlimited fill: exact LU
    L = sp.csr_matrix((f["lv"], f["li"], f["lp"]), shape=(n, n))
    U = sp.csr_matrix((f["uv"], f["ui"], f["up"]), shape=(n, n))
    assert abs(L @ U - A).max() < 1e-10
    f = bg.host_ilut(rp, ci, v, 1e-3, 10)
    pc = oracle.Precond(oracle.PC_ILU0, n, "f64", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
    it_ilut = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 500, 1e-8, pc)[2]
    assert it_ilut < g["f64_meta"][1]


def itsol_iluk(rp, ci, v, level):
    """ILU(level) of the ITSOL code the reference vendors (src/iluk.cpp ilukC / lofC, compiled unchanged into
    oracle/_ref/libref_iluk.so): per row the strictly-lower and strictly-upper columns and float values, and 1/diag"""
    L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_iluk.so"))
    n = len(rp) - 1
    rp32, ci32, v32 = np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32), np.ascontiguousarray(v, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.refk_factor(n, p(rp32), p(ci32), p(v32), int(level)) == 0
    nl, nu = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert L.refk_row_counts(p(nl), p(nu)) == 0
    lc, lv = np.zeros(nl.sum(), np.int32), np.zeros(nl.sum(), np.float32)
    uc, uv = np.zeros(nu.sum(), np.int32), np.zeros(nu.sum(), np.float32)
    dinv = np.zeros(n, np.float32)
    assert L.refk_get(p(lc), p(lv), p(uc), p(uv), p(dinv)) == 0
    return nl, lc, lv, nu, uc, uv, dinv


REF_ILUK = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_iluk.so")


@pytest.mark.skipif(not os.path.exists(REF_ILUK), reason="oracle/_ref/libref_iluk.so not built (needs /root/reference)")
@pytest.mark.parametrize("case", ["pg2d", "th3d", "random"])
def test_iluk_pattern_equals_vendored_itsol(bg, case):
    """bg_host_iluk against the ILU(k) the reference vendors: fill pattern identical for every level (integer
    artefact: bit-exact), values within float round-off (ITSOL multiplies by a stored reciprocal pivot, the product
    divides), level 0 identical with bg_host_ilu0, L U = A on the pattern."""
    from b200gmres import synth

    if case == "pg2d":
        rp, ci, v, _ = synth.pg2d(21, 17, seed=11)
    elif case == "th3d":
        rp, ci, v, _ = synth.th3d(7, 6, 5, seed=12)
    else:  # unstructured, non-symmetric pattern, diagonally dominant
        rng = np.random.default_rng(3)
        n = 300
        M = sp.random(n, n, density=0.012, random_state=7, format="csr") + sp.random(n, n, density=0.004, random_state=8, format="csr").T
        M = sp.csr_matrix(M)
        M.data = rng.uniform(-1, 1, size=M.nnz)
        M = sp.csr_matrix(M + sp.diags(np.asarray(abs(M).sum(axis=1)).ravel() + 1.0))
        M.sort_indices()
        rp, ci, v = M.indptr, M.indices, M.data
    n = len(rp) - 1
    v = np.asarray(v, np.float64)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    f0 = bg.host_ilu0(rp, ci, v)
    k0 = bg.host_iluk(rp, ci, v, 0)
    assert all(np.array_equal(f0[k], k0[k]) for k in ("lp", "li", "lv", "up", "ui", "uv"))
    prev_nnz = 0
    for level in (0, 1, 2, 3, 5):
        f = bg.host_iluk(rp, ci, v.astype(np.float32), level)          # ITSOL's precision
        nl, lc, lv, nu, uc, uv, dinv = itsol_iluk(rp, ci, v, level)
        lp, up = f["lp"], f["up"]      # L carries its unit diagonal last, U its diagonal first
        a = b = 0
        dropped = 0
        for i in range(n):
            for mine_c, mine_v, ref_c, ref_v in ((f["li"][lp[i]:lp[i + 1] - 1], f["lv"][lp[i]:lp[i + 1] - 1], lc[a:a + nl[i]], lv[a:a + nl[i]]),
                                                 (f["ui"][up[i] + 1:up[i + 1]], f["uv"][up[i] + 1:up[i + 1]], uc[b:b + nu[i]], uv[b:b + nu[i]])):
                order = np.argsort(ref_c, kind="stable")        # ITSOL keeps U rows in discovery order
                ref_c, ref_v = ref_c[order], ref_v[order]
                if len(mine_c) != len(ref_c):
                    # the library's structural-zero rule (|v| < 1e-9, SURVEY Q5, splitLU_csr) removes vanishing fill
                    keep = np.isin(ref_c, mine_c)
                    assert np.all(np.abs(ref_v[~keep]) < 2e-9), (level, i)
                    dropped += int((~keep).sum())
                    ref_c, ref_v = ref_c[keep], ref_v[keep]
                assert np.array_equal(mine_c, ref_c), (level, i)
                assert np.allclose(mine_v, ref_v, rtol=2e-5, atol=1e-6)
            assert np.isclose(1.0 / f["uv"][up[i]], dinv[i], rtol=2e-5)
            a += nl[i]; b += nu[i]
        assert dropped == 0 or case == "random"      # the grids have no vanishing fill: their patterns are lofC's exactly
        # FP64 factors on the same pattern: the product reproduces A wherever the factors store an entry
        d = bg.host_iluk(rp, ci, v, level)
        assert dropped or (np.array_equal(d["li"], f["li"]) and np.array_equal(d["ui"], f["ui"]))
        L = sp.csr_matrix((d["lv"], d["li"], d["lp"]), shape=(n, n))
        U = sp.csr_matrix((d["uv"], d["ui"], d["up"]), shape=(n, n))
        P = sp.csr_matrix(abs(L) + abs(U)); P.data[:] = 1
        assert abs((L @ U - A).multiply(P)).max() <= (1e-8 if dropped else 1e-12) * abs(A).max()   # dropped fill: |v| < 1e-9 each
        assert L.nnz + U.nnz >= prev_nnz
        prev_nnz = L.nnz + U.nnz


@pytest.mark.skipif(not os.path.exists(REF_ILUK), reason="oracle/_ref/libref_iluk.so not built (needs /root/reference)")
def test_iluk_fuzz_against_vendored_itsol(bg):
    """40 random sparse matrices (non-symmetric patterns, random row lengths, well-scaled values so that no fill
    vanishes): strictly-lower and strictly-upper patterns of every row identical with lofC's for levels 0..3"""
    for seed in range(40):
        rng = np.random.default_rng(500 + seed)
        n = int(rng.integers(20, 160))
        M = sp.random(n, n, density=float(rng.uniform(0.02, 0.08)), random_state=int(seed), format="csr")
        M.data = rng.uniform(0.5, 1.5, size=M.nnz) * rng.choice([-1.0, 1.0], size=M.nnz)
        M = sp.csr_matrix(M + sp.diags(np.asarray(abs(M).sum(axis=1)).ravel() + 1.0))
        M.sort_indices()
        rp, ci, v = M.indptr, M.indices, M.data
        for level in range(4):
            f = bg.host_iluk(rp, ci, v, level)
            nl, lc, lv, nu, uc, uv, dinv = itsol_iluk(rp, ci, v, level)
            lp, up = f["lp"], f["up"]
            small = min(np.abs(lv).min() if len(lv) else 1.0, np.abs(uv).min() if len(uv) else 1.0)
            if small < 1e-8:
                continue        # a vanishing fill entry: the structural-zero rule would remove it (covered above)
            assert np.array_equal(np.diff(lp) - 1, nl) and np.array_equal(np.diff(up) - 1, nu), (seed, level)
            a = b = 0
            for i in range(n):
                assert np.array_equal(f["li"][lp[i]:lp[i + 1] - 1], lc[a:a + nl[i]]), (seed, level, i)
                assert np.array_equal(f["ui"][up[i] + 1:up[i + 1]], np.sort(uc[b:b + nu[i]])), (seed, level, i)
                a += nl[i]; b += nu[i]


@pytest.mark.parametrize("name", ["iluk_pg2d_13x11", "iluk_th3d_6x5x4"])
def test_iluk_equals_committed_itsol_output(bg, name):
    """the same pin against the committed output of the vendored ITSOL code (tests/golden/make_frontend_golden.py)"""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    rp, ci, v = g["rowptr"], g["col"], g["val"]
    n = len(rp) - 1
    for level in range(4):
        f = bg.host_iluk(rp, ci, v.astype(np.float32), level)
        nl, lc, lv, nu, uc, uv, dinv = (g[f"k{level}_{k}"] for k in ("nl", "lc", "lv", "nu", "uc", "uv", "dinv"))
        lp, up = f["lp"], f["up"]
        assert np.array_equal(np.diff(lp) - 1, nl) and np.array_equal(np.diff(up) - 1, nu)
        a = b = 0
        for i in range(n):
            order = np.argsort(uc[b:b + nu[i]], kind="stable")
            assert np.array_equal(f["li"][lp[i]:lp[i + 1] - 1], lc[a:a + nl[i]])
            assert np.array_equal(f["ui"][up[i] + 1:up[i + 1]], uc[b:b + nu[i]][order])
            assert np.allclose(f["lv"][lp[i]:lp[i + 1] - 1], lv[a:a + nl[i]], rtol=2e-5, atol=1e-6)
            assert np.allclose(f["uv"][up[i] + 1:up[i + 1]], uv[b:b + nu[i]][order], rtol=2e-5, atol=1e-6)
            assert np.isclose(1.0 / f["uv"][up[i]], dinv[i], rtol=2e-5)
            a += nl[i]; b += nu[i]


def test_iluk_improves_convergence_and_respects_blocks(bg, oracle):
    g = load_golden("pg2d_24x20")
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    its = []
    for level in (0, 1, 2, 4):
        f = bg.host_iluk(rp, ci, v, level)
        pc = oracle.Precond(oracle.PC_ILU0, n, "f64", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
        its.append(oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 500, 1e-8, pc)[2])
    assert its[0] == g["f64_meta"][1] and its[0] > its[1] > its[2] > its[3]
    blk = (np.arange(n) * 7 // n).astype(np.int32)
    f = bg.host_iluk(rp, ci, v, 2, blk)
    rows = np.repeat(np.arange(n), np.diff(f["lp"]))
    assert np.all(blk[rows] == blk[f["li"]])
    rows = np.repeat(np.arange(n), np.diff(f["up"]))
    assert np.all(blk[rows] == blk[f["ui"]])
    with pytest.raises(bg.BgError):
        bg.host_iluk(rp, ci, v, -1)


def test_level_stream_packers(tmp_path):
    """The two BG_TRSV_LEVEL streams (sync-free slices for any factor, index-free strips for 5-point ILU(0)),
    interpreted on the host the way trsv_syncfree / trsv_lattice_global consume them: no row is read before it
    is published (no device deadlock), slices never span levels, explicit |v| < 1e-9 entries are scheduled
    structurally, and the result equals a plain CSR forward/backward solve."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "stream_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "cpp", "stream_check.cpp"),
                           os.path.join(root, "gpu-gmres_b200", "host", "host_algos.cpp"),
                           "/usr/local/cuda/lib64/libmetis_static.a", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_sell_packer(tmp_path):
    """build_sell (the SpMV storage of every matrix) at single-thread and threaded sizes: entries in row order
    at [slice_ptr + k*slice + t], zero padding with in-range columns, product identical to CSR."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "sell_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "cpp", "sell_check.cpp"),
                           os.path.join(root, "gpu-gmres_b200", "host", "host_algos.cpp"),
                           "/usr/local/cuda/lib64/libmetis_static.a", "-lpthread", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_fused_block_stream_packer(tmp_path):
    """The packed factor stream of the block-Jacobi solve, interpreted on the host step by step the
    way trsv_block_fused consumes it, reproduces a plain CSR forward/backward solve."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "pack_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "cpp", "pack_check.cpp"),
                           os.path.join(root, "gpu-gmres_b200", "host", "host_algos.cpp"),
                           "/usr/local/cuda/lib64/libmetis_static.a", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
