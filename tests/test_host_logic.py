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
