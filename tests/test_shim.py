"""The reference-facing C++ shim (gpu-gmres_b200/host/shim): the reference's own driver flows
(thermal main1/main2, power-grid gmresInterfacePGfloat) written against the reference's header
names, run on the GPU through libb200gmres_shim.so and compared with the oracle / golden vectors."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gpu-gmres_b200")


def build_driver(tmp_path):
    exe = str(tmp_path / "shim_driver")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "host", "shim"),
                           "-I" + os.path.join(PKG, "host", "shim", "standalone"),
                           "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"), "-L" + PKG,
                           "-lb200gmres_shim", "-lb200gmres", "-Wl,-rpath," + PKG, "-o", exe])
    return exe


def write_case(path, rp, ci, v, b, extra=b""):
    n = len(rp) - 1
    with open(path, "wb") as f:
        f.write(struct.pack("ii", n, len(ci)))
        f.write(np.asarray(rp, np.int32).tobytes())
        f.write(np.asarray(ci, np.int32).tobytes())
        f.write(np.asarray(v, np.float32).tobytes())
        f.write(np.asarray(b, np.float32).tobytes())
        f.write(extra)


def read_records(path, n):
    raw = open(path, "rb").read()
    rec, out = 12 + 4 * n, []
    for k in range(len(raw) // rec):
        flag, iters, tol = struct.unpack("iif", raw[k * rec:k * rec + 12])
        out.append((flag, iters, tol, np.frombuffer(raw[k * rec + 12:(k + 1) * rec], np.float32).astype(np.float64)))
    return out


def test_shim_library_exports_reference_api():
    """no GPU needed: the shim library exists and defines the reference's entry points"""
    lib = os.path.join(PKG, "libb200gmres_shim.so")
    assert os.path.exists(lib), "run __graft_entry__.build()"
    syms = subprocess.run(["nm", "-DC", "--defined-only", lib], capture_output=True, text=True).stdout
    for name in ("GMRES_GPU(", "GMRES_GPU_tran(", "GMRESilu_GPU(", "sgemv_GPU(", "getLevel(", "leftILU(", "generateLevel(",
                 "gmresInterfacePGfloat::setPrecondPG(", "gmresInterfacePGfloat::GMRES_dev_PG(", "MyILU0::Initilize(",
                 "MyILUPPfloat::Initilize(", "MyDIAG::Initilize(", "MyNONE::Initilize(", "Preconditioner::DevPrecond_left(",
                 "Preconditioner::DevPrecond_right(", "Preconditioner::DevPrecond_starting_value(", "MySpMatrix::Initilize(",
                 "allocateSparseMatrixGPU(", "gpuMallocCpyCSR(", "freeSparseMatrixGPU(", "mySpMatrixFree(", "genCSRFormat(",
                 "GMRES_GPU_Data::Initilize(", "setGPUdevice(",
                 # round 2: the rest of the surface the unchanged drivers pull (VERDICT r1 "close the boundary")
                 "gmresInterfacePG::setPrecondPG(", "gmresInterfacePG::GMRES_host_PG(", "gmresInterfacePGfloat::GMRES_host_PG(",
                 "MyILUPP::Initilize(", "MyAINV::Initilize(", "Update_GPU(", "GMRES_GPU_leftDiag(", "GMRES_GPU_leftILU0(",
                 "GMRES(float const*", "GMRESilu(float const*", "GMRES_tran(float const*", "GMRES_leftILU0(", "GMRES_leftDiag(",
                 "csr2csc(", "splitLU_csc(", "splitLU_csr(", "Preconditioner::HostPrecond(", "ApplyPlaneRotation(",
                 "GeneratePlaneRotation(", "difftime(", "SpMV(float*", "SpMV_s(float*"):
        assert name in syms, name


REF = "/root/reference"
BUILD = os.path.join(ROOT, "tests", "_build")  # git-ignored, travels to the GPU box (the reference tree does not)


def _nvcc(args):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-w"] + args)


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_reference_thermal_drivers_compile_and_link_unchanged(tmp_path):
    """src_thermal/main1.cu and main2.cu, UNCHANGED, compile against the shim's headers and link against
    libb200gmres_shim.so (+ the reference's own SpMV_gen/compute/inspect.cpp, which stay in its build).  The files are
    copied to a scratch directory only because a quoted #include looks next to the including file first; the shim
    headers come first on the include path, the tree's own config.h / defs.h / SpMV_inspect.h / rightLookingILU.h
    second, stand-ins for helper_cuda.h / cutil.h / cusp (absent from this image) last."""
    import shutil

    th = os.path.join(REF, "src_thermal")
    inc = ["-I" + os.path.join(PKG, "host", "shim"), "-I" + th, "-I" + os.path.join(ROOT, "oracle", "shim")]
    objs = []
    for f in ("SpMV_gen", "SpMV_compute", "SpMV_inspect"):
        subprocess.check_call(["g++", "-O2", "-w", "-c", os.path.join(th, f + ".cpp"), "-o", str(tmp_path / (f + ".o"))])
        objs.append(str(tmp_path / (f + ".o")))
    os.makedirs(BUILD, exist_ok=True)
    for main, exe in (("main1.cu", "ref_gmres_SpMV"), ("main2.cu", "ref_gmres_tran")):
        shutil.copy(os.path.join(th, main), tmp_path / main)
        os.chmod(tmp_path / main, 0o644)
        _nvcc(inc + ["-c", str(tmp_path / main), "-o", str(tmp_path / (main + ".o"))])
        # --cudart shared: the driver launches the shim's SpMV kernel symbol itself, so both must use ONE CUDA runtime
        _nvcc(["--cudart", "shared", "-o", os.path.join(BUILD, exe), str(tmp_path / (main + ".o"))] + objs +
              ["-L" + PKG, "-lb200gmres_shim", "-lb200gmres", "-lcublas", "-Xlinker", "-rpath," + PKG, "-Xlinker", "-rpath,/usr/local/cuda/lib64"])
        assert os.path.exists(os.path.join(BUILD, exe))


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_reference_power_grid_call_sites_compile_unchanged(tmp_path):
    """The gmresInterfacePG / gmresInterfacePGfloat call sites of src/mna_solve_gpu_gmres.cpp (507-532, 608-621),
    extracted verbatim at test time, compile against the shim's gmres_interface_pg.h and link."""
    lines = open(os.path.join(REF, "src", "mna_solve_gpu_gmres.cpp"), errors="replace").read().split("\n")
    a, b = lines[506:532], lines[607:621]
    assert "gmresInterfacePG GmyInterfacePG" in a[0] and "GMRES_dev_PG" in a[-2] and "useDoubleILU" in b[0] and b[-1].strip() == "}"
    (tmp_path / "pg_callsite_A.inc").write_text("\n".join(a) + "\n")
    (tmp_path / "pg_callsite_B.inc").write_text("\n".join(b) + "\n")
    (tmp_path / "main.cpp").write_text("int main() { return 0; }\n")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-I" + str(tmp_path), "-I" + os.path.join(PKG, "host", "shim"),
                           "-I" + os.path.join(PKG, "host", "shim", "standalone"), "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "cpp", "pg_callsite_wrap.cpp"), str(tmp_path / "main.cpp"),
                           "-L" + PKG, "-lb200gmres_shim", "-lb200gmres", "-Wl,-rpath," + PKG, "-o", str(tmp_path / "pg_callsite")])


@pytest.mark.gpu
def test_reference_thermal_driver_runs_unchanged(tmp_path, oracle):
    """The reference's own gmres_SpMV (src_thermal/main1.cu, built unchanged against the shim by the CPU test above in the
    build container) on a small thermal grid: its SpMV<<<>>> self-check passes, and both the host-array GMRES() call
    (main1.cu:372, run on the device here) and GMRES_GPU (main1.cu:396) reproduce the oracle's FP32 solve."""
    exe = os.path.join(BUILD, "ref_gmres_SpMV")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/ref_gmres_SpMV was not built (needs the reference tree)")
    g = load_golden("th3d_8x7x6")
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    with open(tmp_path / "A.mtx", "w") as f:
        f.write(f"{n} {n} {len(ci)}\n")
        for r in range(n):
            for p in range(rp[r], rp[r + 1]):
                f.write(f"{r + 1} {ci[p] + 1} {float(np.float32(v[p])):.9e}\n")
    with open(tmp_path / "b.txt", "w") as f:
        for e in b:
            f.write(f"{float(np.float32(e)):.9e}\n")
    out = subprocess.run([exe, "A.mtx", "b.txt"], input="2\n", capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Test PASSED" in out.stdout and "GPU GMRES flag = 0" in out.stdout and "CPU GMRES flag = 0" in out.stdout, out.stdout[-2000:]
    for name in ("xGPU.txt", "xCPU.txt"):
        x = np.loadtxt(tmp_path / name)
        assert np.linalg.norm(x - g["f64_x"]) <= 5e-4 * np.linalg.norm(g["f64_x"]), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pg2d_24x20", "th3d_8x7x6"])
def test_thermal_driver_flow_matches_oracle(tmp_path, oracle, name):
    g = load_golden(name)
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    exe = build_driver(tmp_path)
    write_case(tmp_path / "in.bin", rp, ci, v, b)
    # MyILU0 + GMRES_GPU (src_thermal/main1.cu:335-397), FP32 like the reference
    subprocess.check_call([exe, str(tmp_path / "in.bin"), str(tmp_path / "o1.bin"), "thermal-ilu0"])
    (flag, iters, tol, x), (_, _, _, z) = read_records(tmp_path / "o1.bin", n)
    f = oracle.ilu0(rp, ci, v, "f32")
    pc = oracle.Precond(oracle.PC_ILU0, n, "f32", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
    oflag, xo, oit, ores, hist = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 2000, 1e-6, pc, hist_cap=2000)
    assert flag == 0 and oflag == 0 and tol < 1e-6
    first64 = int(np.argmax(g["f64_hist"] < 1e-6)) + 1
    assert first64 - 3 <= iters <= oit + 3
    assert np.linalg.norm(x - g["f64_x"]) <= 5e-4 * np.linalg.norm(g["f64_x"])
    want = pc.apply(0, b.astype(np.float32))
    assert np.linalg.norm(z - want) <= 5e-6 * np.linalg.norm(want)  # DevPrecond on its own
    # MyDIAG and MyNONE
    for mode, key in (("thermal-diag", "diag"), ("thermal-none", "none")):
        subprocess.check_call([exe, str(tmp_path / "in.bin"), str(tmp_path / "o2.bin"), mode])
        flag, iters, tol, x = read_records(tmp_path / "o2.bin", n)[0]
        if g[key + "_meta"][0] == 0:  # the FP64 oracle converged at 1e-8; the FP32 driver stops at 1e-6
            assert flag == 0 and np.linalg.norm(x - g[key + "_x"]) <= 1e-3 * np.linalg.norm(g[key + "_x"])
    # transient pattern (main2.cu): workspace reuse + warm start => later steps need fewer iterations
    subprocess.check_call([exe, str(tmp_path / "in.bin"), str(tmp_path / "o3.bin"), "tran"])
    recs = read_records(tmp_path / "o3.bin", n)
    assert [r[0] for r in recs] == [0, 0, 0]
    import scipy.sparse as sp
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    for k, r in enumerate(recs):
        bk = b * (1 + 0.05 * k)
        assert np.linalg.norm(A @ r[3] - bk) <= 2e-4 * np.linalg.norm(bk)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pg2d_24x20", "th3d_8x7x6"])
def test_power_grid_interface_matches_golden(tmp_path, name):
    """gmresInterfacePGfloat::setPrecondPG + GMRES_dev_PG on the reference's ILU++ factors."""
    g = load_golden(name + "_split")
    rp, ci, v, b = g["rowptr"], g["col"], g["val"], g["b"]
    n = len(rp) - 1
    extra = struct.pack("ii", len(g["li"]), len(g["ui"]))
    for a, dt in ((g["lp"], np.int32), (g["li"], np.int32), (g["lv"], np.float64), (g["up"], np.int32), (g["ui"], np.int32),
                  (g["uv"], np.float64), (g["middle"], np.float32), (g["lscale"], np.float32), (g["rscale"], np.float32),
                  (g["permRow"], np.int32), (g["permCol"], np.int32)):
        extra += np.asarray(a, dt).tobytes()
    exe = build_driver(tmp_path)
    write_case(tmp_path / "in.bin", rp, ci, v, b, extra)
    subprocess.check_call([exe, str(tmp_path / "in.bin"), str(tmp_path / "o.bin"), "pg"])
    (f0, it0, tol0, x0), (f1, it1, tol1, x1) = read_records(tmp_path / "o.bin", n)
    assert f0 == 0 and f1 == 0 and tol0 < 1e-7 and tol1 < 1e-7
    assert np.linalg.norm(x0 - g["f64_x"]) <= 2e-4 * np.linalg.norm(g["f64_x"])
    # FP32 vectors at tol 1e-7 sit at the FP32 round-off floor: the FP32 oracle (float dot accumulators)
    # stagnates there, the device accumulates dots in FP64.  Bounds: not before the FP64 trajectory
    # crosses the tolerance, not after the FP32 oracle.
    first64 = int(np.argmax(g["f64_hist"] < 1e-7)) + 1
    assert first64 - 3 <= it0 <= int(g["f32f64_meta"][1]) + 3
    assert np.linalg.norm(x1 - 1.01 * g["f64_x"]) <= 3e-4 * np.linalg.norm(g["f64_x"])
    assert it1 < it0  # warm start from the DC solution
