"""Matrix / vector files in the reference's own formats: bg_host_read_matrix & co. against the
reference's readers and writers compiled unchanged into oracle/_ref (src_thermal/SpMV_gen.cpp readSparseMatrix +
genCSRFormat, readInputVector, writeOutputVector; src/formatConvert.cpp writeCSR)."""
import numpy as np
import pytest


def _triplet_file(path, rp, ci, v, seed):
    rng = np.random.default_rng(seed)
    n = len(rp) - 1
    rows = np.repeat(np.arange(n), np.diff(rp))
    order = rng.permutation(len(ci))  # the readers sort by (row, col): any file order is legal
    with open(path, "w") as f:
        f.write("%%MatrixMarket-like triplets written by the test\n% second comment line\n")
        f.write(f"{n} {n} {len(ci)}\n")
        for k in order:
            f.write(f"{rows[k] + 1} {ci[k] + 1} {v[k]:.9e}\n")


@pytest.mark.parametrize("case", ["pg2d", "th3d", "rand"])
def test_triplet_reader_matches_reference(bg, oracle, tmp_path, case):
    from b200gmres import synth

    rp, ci, v, _ = {"pg2d": lambda: synth.pg2d(13, 9, seed=3), "th3d": lambda: synth.th3d(5, 4, 3, seed=4),
                    "rand": lambda: synth.random_sparse(120, 6, seed=5)}[case]()
    path = tmp_path / "A.mtx"
    _triplet_file(path, rp, ci, v, seed=7)
    rp2, ci2, v2, ncols, empty = bg.host_read_matrix(path, bg.FILE_TRIPLETS)
    assert ncols == len(rp) - 1 and empty == 0 and v2.dtype == np.float32
    assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci)
    assert np.array_equal(v2, np.array([np.float32(f"{x:.9e}") for x in v]))
    if oracle.ref_io_lib() is None:
        pytest.skip("oracle/_ref/libref_io.so not built")
    rr, rc, rv, rn = oracle.ref_read_sparse_matrix(path)
    assert rn == ncols and np.array_equal(rr, rp2) and np.array_equal(rc, ci2) and np.array_equal(rv, v2)  # bit-exact


def test_triplet_reader_errors_and_empty_rows(bg, tmp_path):
    p = tmp_path / "e.mtx"
    p.write_text("% c\n4 4 3\n1 1 2.0\n4 2 -1.5\n4 4 3.0\n")  # rows 2 and 3 have no entry
    rp, ci, v, ncols, empty = bg.host_read_matrix(p, bg.FILE_TRIPLETS)
    assert list(rp) == [0, 1, 1, 1, 3] and list(ci) == [0, 1, 3] and empty == 2 and list(v) == [2.0, -1.5, 3.0]
    p.write_text("3 3 2\n1 1 1.0\n4 1 1.0\n")
    with pytest.raises(Exception):
        bg.host_read_matrix(p, bg.FILE_TRIPLETS)
    p.write_text("3 3 2\n1 1 1.0\n")
    with pytest.raises(Exception):
        bg.host_read_matrix(p, bg.FILE_TRIPLETS)
    with pytest.raises(Exception):
        bg.host_read_matrix(tmp_path / "missing.mtx", bg.FILE_TRIPLETS)


def test_binary_csr_round_trip_and_reference_writer(bg, oracle, tmp_path):
    from b200gmres import synth

    rp, ci, v, _ = synth.pg2d(17, 11, seed=9)
    n = len(rp) - 1
    for fmt, dt in ((bg.FILE_CSR_F32, np.float32), (bg.FILE_CSR_F64, np.float64)):
        path = tmp_path / f"A{fmt}.bin"
        bg.host_write_csr(path, fmt, rp, ci, v)  # values converted to the file's type
        rp2, ci2, v2, ncols, _ = bg.host_read_matrix(path, fmt)
        assert ncols == n and v2.dtype == dt and np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and np.array_equal(v2, v.astype(dt))
    if oracle.ref() is None:
        pytest.skip("oracle/_ref/libref.so not built")
    # a file written by the reference's writeCSR (src/formatConvert.cpp:425-445) reads back bit-exactly, and our
    # writer produces the same bytes
    pr, pb = tmp_path / "ref.bin", tmp_path / "bg.bin"
    oracle.ref_write_csr(pr, rp, ci, v, n)
    bg.host_write_csr(pb, bg.FILE_CSR_F32, rp, ci, v.astype(np.float32))
    assert pr.read_bytes() == pb.read_bytes()
    rp3, ci3, v3, _, _ = bg.host_read_matrix(pr, bg.FILE_CSR_F32)
    assert np.array_equal(rp3, rp) and np.array_equal(ci3, ci) and np.array_equal(v3, v.astype(np.float32))
    bad = tmp_path / "short.bin"
    bad.write_bytes(pr.read_bytes()[:-8])
    with pytest.raises(Exception):
        bg.host_read_matrix(bad, bg.FILE_CSR_F32)


def test_vector_files_match_reference(bg, oracle, tmp_path):
    x = np.random.default_rng(1).standard_normal(37)
    p = tmp_path / "x.txt"
    bg.host_write_vector(p, x)
    assert np.array_equal(bg.host_read_vector(p, 37, bg.F64), x)  # "%20.18e" round-trips a double
    with pytest.raises(Exception):
        bg.host_read_vector(p, 38, bg.F64)
    if oracle.ref_io_lib() is None:
        pytest.skip("oracle/_ref/libref_io.so not built")
    xf = x.astype(np.float32)
    pr = tmp_path / "xr.txt"
    oracle.ref_write_output_vector(pr, xf)         # the reference's writer ...
    pb = tmp_path / "xb.txt"
    bg.host_write_vector(pb, xf)                   # ... and ours produce the same text
    assert pr.read_text() == pb.read_text()
    assert np.array_equal(bg.host_read_vector(pr, 37, bg.F32), xf)
    assert np.array_equal(oracle.ref_read_input_vector(pb, 37), xf)


@pytest.mark.gpu
def test_file_driver_end_to_end(bg, tmp_path):
    """tools/gmres_file.py = the thermal driver's flow (src_thermal/main1.cu) on files in the reference's formats."""
    import os
    import subprocess
    import sys

    import scipy.sparse as sp
    from b200gmres import synth

    rp, ci, v, b = synth.th3d(9, 8, 7, seed=21)
    n = len(rp) - 1
    pa, pb, px = tmp_path / "A.mtx", tmp_path / "b.txt", tmp_path / "x.txt"
    _triplet_file(pa, rp, ci, v, seed=2)
    bg.host_write_vector(pb, b.astype(np.float32))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ([], ["--dtype", "f64", "--tol", "1e-8", "--schedule", "block"]):
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "gmres_file.py"), str(pa), str(pb), str(px)] + extra,
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "GPU GMRES flag = 0" in out.stdout
        x = bg.host_read_vector(px, n, bg.F64)
        A = sp.csr_matrix((v, ci, rp), shape=(n, n))
        rel = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
        assert rel < (1e-5 if extra else 2e-4)  # the rhs file carries float32 digits; the stopping test is on the preconditioned residual
