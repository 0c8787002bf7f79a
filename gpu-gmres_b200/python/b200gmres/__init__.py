"""ctypes binding of libb200gmres.so (the C ABI in include/b200gmres.h) plus synthetic workloads.

Python is only the test/bench harness here; the product is the CUDA library and the C++ shim
that mirrors the reference's solver API (gpu-gmres_b200/host/shim).  Nothing in this module
computes: every call goes through the C ABI, and a missing library is a hard error.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import synth  # noqa: F401

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # gpu-gmres_b200/
LIB_PATH = os.path.join(_PKG, "libb200gmres.so")
HEADER = os.path.join(os.path.dirname(_PKG), "include", "b200gmres.h")

F32, F64 = 0, 1
HOST, DEVICE = 0, 1
SPMV_AUTO, SPMV_CSR_VECTOR, SPMV_SELL, SPMV_CSR_WARP = 0, 1, 2, 3
PC_FULL, PC_LEFT, PC_RIGHT, PC_START, PC_RHS = 0, 1, 2, 3, 4
ORTH_MGS, ORTH_CGS_ADAPTIVE, ORTH_CGS2 = 0, 1, 2
TRSV_LEVEL, TRSV_BLOCK = 0, 1
PART_CONTIGUOUS, PART_METIS, PART_GRID2D, PART_GRID3D = 0, 1, 2, 3
ERR_NO_GPU = -3

_NP = {F32: np.float32, F64: np.float64}


def dtype_code(dt) -> int:
    return F32 if np.dtype(dt) == np.float32 else F64


class BgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200gmres error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile libb200gmres.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _PKG, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class MatrixInfo(C.Structure):
    _fields_ = [("n", C.c_int), ("nnz", C.c_int64), ("dtype", C.c_int), ("format", C.c_int),
                ("max_row_nnz", C.c_int), ("mean_row_nnz", C.c_double), ("hist", C.c_int64 * 8),
                ("sell_padded_nnz", C.c_int64), ("sell_slice", C.c_int), ("spmv_bytes", C.c_int64)]


class FactorsInfo(C.Structure):
    _fields_ = [("n", C.c_int), ("dtype", C.c_int), ("nnzL", C.c_int64), ("nnzU", C.c_int64), ("zero_pivots", C.c_int)]


class PrecondOpts(C.Structure):
    _fields_ = [("trsv_schedule", C.c_int), ("block_method", C.c_int), ("block_rows", C.c_int), ("dims", C.c_int * 6),
                ("factor_dtype", C.c_int), ("use_tma", C.c_int), ("ilut_lfil", C.c_int), ("ilut_tau", C.c_double),
                ("chunk_bytes", C.c_int), ("stages", C.c_int), ("warps_per_cta", C.c_int), ("lattice", C.c_int), ("device_factor", C.c_int)]


class PrecondInfo(C.Structure):
    _fields_ = [("kind", C.c_int), ("n", C.c_int), ("dtype", C.c_int), ("factor_dtype", C.c_int), ("trsv_schedule", C.c_int),
                ("nnzL", C.c_int64), ("nnzU", C.c_int64), ("levelsL", C.c_int), ("levelsU", C.c_int),
                ("nblocks", C.c_int), ("max_block_rows", C.c_int), ("trsv_bytes", C.c_int64), ("stream_bytes", C.c_int64),
                ("lattice", C.c_int), ("device_factored", C.c_int)]


class GmresOpts(C.Structure):
    _fields_ = [("m", C.c_int), ("max_iter", C.c_int), ("tol", C.c_double), ("orth", C.c_int), ("reorth_eta", C.c_double),
                ("poll_chunk", C.c_int), ("use_graph", C.c_int), ("history_cap", C.c_int)]


class GmresResult(C.Structure):
    _fields_ = [("flag", C.c_int), ("iters", C.c_int), ("iters_done", C.c_int), ("resid", C.c_double),
                ("restarts", C.c_int), ("reorths", C.c_int), ("launches", C.c_int64), ("solve_ms", C.c_double)]


class PhaseStats(C.Structure):
    _fields_ = [("spmv_ms", C.c_double), ("trsv_ms", C.c_double), ("orth_ms", C.c_double), ("other_ms", C.c_double),
                ("spmv_bytes", C.c_int64), ("trsv_bytes", C.c_int64), ("orth_bytes", C.c_int64), ("other_bytes", C.c_int64),
                ("spmv_calls", C.c_int64), ("trsv_calls", C.c_int64), ("orth_calls", C.c_int64),
                ("dots_ms", C.c_double), ("update_ms", C.c_double), ("dots_bytes", C.c_int64), ("update_bytes", C.c_int64),
                ("dots_calls", C.c_int64), ("update_calls", C.c_int64), ("halo_ms", C.c_double), ("halo_calls", C.c_int64)]


_lib = None


def lib():
    """The loaded C ABI.  Raises if the library is missing (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built: run __graft_entry__.build() (make -C gpu-gmres_b200)")
        L = C.CDLL(LIB_PATH)
        L.bg_last_error.restype = C.c_char_p
        L.bg_context_launch_count.restype = C.c_int64
        _lib = L
    return _lib


def _chk(rc):
    if rc < 0:
        raise BgError(rc, lib().bg_last_error().decode())
    return rc


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw device pointer


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def device_count() -> int:
    return lib().bg_device_count()


# ------------------------------------------------------------------ host-only artefacts
def host_factor_levels(rowptr, col, val):
    n = len(rowptr) - 1
    level = np.zeros(n, np.int32)
    _chk(lib().bg_host_factor_levels(n, _p(_i32(rowptr)), _p(_i32(col)), _p(np.ascontiguousarray(val, np.float32)), _p(level)))
    return level


def host_trsv_levels(is_lower, rowptr, col, val):
    n = len(rowptr) - 1
    val = np.ascontiguousarray(val)
    level = np.zeros(n, np.int32)
    _chk(lib().bg_host_trsv_levels(int(is_lower), n, _p(_i32(rowptr)), _p(_i32(col)), _p(val), dtype_code(val.dtype), _p(level)))
    return level


def host_bucket_levels(level):
    level = _i32(level)
    n = len(level)
    nl = C.c_int(0)
    lp = np.zeros((int(level.max()) + 2) if n else 1, np.int32)
    nodes = np.zeros(n, np.int32)
    _chk(lib().bg_host_bucket_levels(n, _p(level), C.byref(nl), _p(lp), _p(nodes)))
    return nl.value, lp, nodes


def _factors_out(h):
    info = FactorsInfo()
    _chk(lib().bg_factors_get_info(h, C.byref(info)))
    dt = _NP[info.dtype]
    n = info.n
    r = dict(lp=np.zeros(n + 1, np.int32), li=np.zeros(info.nnzL, np.int32), lv=np.zeros(info.nnzL, dt),
             up=np.zeros(n + 1, np.int32), ui=np.zeros(info.nnzU, np.int32), uv=np.zeros(info.nnzU, dt),
             zero_pivots=info.zero_pivots)
    _chk(lib().bg_factors_get(h, _p(r["lp"]), _p(r["li"]), _p(r["lv"]), _p(r["up"]), _p(r["ui"]), _p(r["uv"])))
    lib().bg_factors_destroy(h)
    return r


def host_ilu0(rowptr, col, val, block_of_row=None):
    n = len(rowptr) - 1
    val = np.ascontiguousarray(val)
    h = C.c_void_p()
    blk = None if block_of_row is None else _i32(block_of_row)
    _chk(lib().bg_host_ilu0(n, _p(_i32(rowptr)), _p(_i32(col)), _p(val), dtype_code(val.dtype), _p(blk), C.byref(h)))
    return _factors_out(h)


def host_ilut(rowptr, col, val, tau, lfil, block_of_row=None):
    n = len(rowptr) - 1
    val = np.ascontiguousarray(val)
    h = C.c_void_p()
    blk = None if block_of_row is None else _i32(block_of_row)
    _chk(lib().bg_host_ilut(n, _p(_i32(rowptr)), _p(_i32(col)), _p(val), dtype_code(val.dtype), C.c_double(tau), int(lfil),
                            _p(blk), C.byref(h)))
    return _factors_out(h)


def host_iluk(rowptr, col, val, level, block_of_row=None):
    """ILU(level) on the host: level-of-fill rule of ITSOL's lofC (the reference's vendored ILUK); level 0 is host_ilu0"""
    n = len(rowptr) - 1
    val = np.ascontiguousarray(val)
    h = C.c_void_p()
    blk = None if block_of_row is None else _i32(block_of_row)
    _chk(lib().bg_host_iluk(n, _p(_i32(rowptr)), _p(_i32(col)), _p(val), dtype_code(val.dtype), int(level), _p(blk), C.byref(h)))
    return _factors_out(h)


FILE_TRIPLETS, FILE_CSR_F32, FILE_CSR_F64 = 0, 1, 2


class HostCsrInfo(C.Structure):
    _fields_ = [("n_rows", C.c_int), ("n_cols", C.c_int), ("nnz", C.c_int64), ("dtype", C.c_int), ("empty_rows", C.c_int)]


def host_read_matrix(path, file_format):
    """Matrix file in one of the reference's formats -> (rowptr, col, val, n_cols, empty_rows)."""
    h = C.c_void_p()
    _chk(lib().bg_host_read_matrix(str(path).encode(), int(file_format), C.byref(h)))
    info = HostCsrInfo()
    _chk(lib().bg_host_csr_get_info(h, C.byref(info)))
    rp = np.zeros(info.n_rows + 1, np.int32); ci = np.zeros(info.nnz, np.int32); v = np.zeros(info.nnz, _NP[info.dtype])
    _chk(lib().bg_host_csr_get(h, _p(rp), _p(ci), _p(v)))
    lib().bg_host_csr_destroy(h)
    return rp, ci, v, info.n_cols, info.empty_rows


def host_write_csr(path, file_format, rowptr, col, val, n_cols=None):
    n = len(rowptr) - 1
    val = np.ascontiguousarray(val)
    _chk(lib().bg_host_write_csr(str(path).encode(), int(file_format), n, n if n_cols is None else int(n_cols), _p(_i32(rowptr)), _p(_i32(col)),
                                 _p(val), dtype_code(val.dtype)))


def host_read_vector(path, n, dtype=F64):
    out = np.zeros(n, _NP[dtype])
    _chk(lib().bg_host_read_vector(str(path).encode(), int(n), int(dtype), _p(out)))
    return out


def host_write_vector(path, x):
    x = np.ascontiguousarray(x)
    _chk(lib().bg_host_write_vector(str(path).encode(), len(x), dtype_code(x.dtype), _p(x)))


def host_partition(Ap, Ai, npart, n_nodes):
    m = len(Ap) - 1
    ps, npt, pinv, q = np.zeros(npart + 1, np.int64), np.zeros(m, np.int64), np.zeros(m, np.int64), np.zeros(m, np.int64)
    _chk(lib().bg_host_partition(C.c_int64(m), _p(_i64(Ap)), _p(_i64(Ai)), npart, C.c_int64(n_nodes), _p(ps), _p(npt), _p(pinv), _p(q)))
    return ps, npt, pinv, q


def host_partition3_post(Ap, Ai, npart, part):
    m = len(Ap) - 1
    ps, npt, pinv, q = np.zeros(npart + 1, np.int64), np.zeros(m, np.int64), np.zeros(m, np.int64), np.zeros(m, np.int64)
    _chk(lib().bg_host_partition3_post(C.c_int64(m), _p(_i64(Ap)), _p(_i64(Ai)), npart, _p(_i64(part)), _p(ps), _p(npt), _p(pinv), _p(q)))
    return ps, npt, pinv, q


def host_block_partition(method, rowptr, col, nblocks, dims=None):
    n = len(rowptr) - 1
    blk = np.zeros(n, np.int32)
    d = None if dims is None else _i32(dims)
    nb = _chk(lib().bg_host_block_partition(method, n, _p(_i32(rowptr)), _p(_i32(col)), nblocks, _p(d), _p(blk)))
    return nb, blk


# ------------------------------------------------------------------ device objects
class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        _chk(lib().bg_context_create(C.byref(self.h), device))
        self.device = device

    def sync(self):
        _chk(lib().bg_context_synchronize(self.h))

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        _chk(lib().bg_context_stream(self.h, C.byref(s)))
        return s.value or 0

    @property
    def launches(self) -> int:
        return lib().bg_context_launch_count(self.h)

    def close(self):
        if self.h:
            lib().bg_context_destroy(self.h)
            self.h = C.c_void_p()


class Matrix:
    def __init__(self, ctx, rowptr, col, val, dtype=F64, fmt=SPMV_AUTO):
        self.ctx = ctx
        rowptr, col = _i32(rowptr), _i32(col)
        val = np.ascontiguousarray(val)
        self.n = len(rowptr) - 1
        self.dtype = dtype
        self.h = C.c_void_p()
        _chk(lib().bg_matrix_create_csr(ctx.h, self.n, C.c_int64(len(col)), _p(rowptr), _p(col), _p(val), dtype_code(val.dtype),
                                        HOST, dtype, fmt, C.byref(self.h)))

    def info(self) -> MatrixInfo:
        i = MatrixInfo()
        _chk(lib().bg_matrix_get_info(self.h, C.byref(i)))
        return i

    def spmv(self, d_x, d_y):
        _chk(lib().bg_spmv(self.ctx.h, self.h, _p(d_x), _p(d_y)))

    def spmv_axpby(self, alpha, d_x, beta, d_b, d_y):
        _chk(lib().bg_spmv_axpby(self.ctx.h, self.h, C.c_double(alpha), _p(d_x), C.c_double(beta), _p(d_b), _p(d_y)))

    def close(self):
        if self.h:
            lib().bg_matrix_destroy(self.h)
            self.h = C.c_void_p()


def precond_opts(**kw) -> PrecondOpts:
    o = PrecondOpts()
    lib().bg_precond_opts_default(C.byref(o))
    for k, v in kw.items():
        if k == "dims":
            for i, d in enumerate(v):
                o.dims[i] = int(d)
        else:
            setattr(o, k, v)
    return o


class Precond:
    def __init__(self, ctx, h, n):
        self.ctx, self.h, self.n = ctx, h, n

    @classmethod
    def none(cls, ctx, n, dtype=F64):
        h = C.c_void_p()
        _chk(lib().bg_precond_create_none(ctx.h, n, dtype, C.byref(h)))
        return cls(ctx, h, n)

    @classmethod
    def diag(cls, ctx, rowptr, col, val, dtype=F64):
        n = len(rowptr) - 1
        val = np.ascontiguousarray(val)
        h = C.c_void_p()
        _chk(lib().bg_precond_create_diag(ctx.h, n, _p(_i32(rowptr)), _p(_i32(col)), _p(val), dtype_code(val.dtype), dtype, C.byref(h)))
        return cls(ctx, h, n)

    @classmethod
    def ilu0(cls, ctx, rowptr, col, val, dtype=F64, opts=None, ilut=False):
        n = len(rowptr) - 1
        val = np.ascontiguousarray(val)
        h = C.c_void_p()
        f = lib().bg_precond_create_ilut if ilut else lib().bg_precond_create_ilu0
        _chk(f(ctx.h, n, _p(_i32(rowptr)), _p(_i32(col)), _p(val), dtype_code(val.dtype), dtype,
               C.byref(opts) if opts is not None else None, C.byref(h)))
        return cls(ctx, h, n)

    @classmethod
    def iluk(cls, ctx, rowptr, col, val, level, dtype=F64, opts=None):
        """ILU(level): level-of-fill factors (bg_host_iluk) behind the same triangular-solve schedules"""
        n = len(rowptr) - 1
        val = np.ascontiguousarray(val)
        h = C.c_void_p()
        _chk(lib().bg_precond_create_iluk(ctx.h, n, _p(_i32(rowptr)), _p(_i32(col)), _p(val), dtype_code(val.dtype), dtype, int(level),
                                          C.byref(opts) if opts is not None else None, C.byref(h)))
        return cls(ctx, h, n)

    @classmethod
    def lu(cls, ctx, lp, li, lv, up, ui, uv, dtype=F64, opts=None):
        n = len(lp) - 1
        lv = np.ascontiguousarray(lv)
        uv = np.ascontiguousarray(uv, dtype=lv.dtype)
        h = C.c_void_p()
        _chk(lib().bg_precond_create_lu(ctx.h, n, _p(_i32(lp)), _p(_i32(li)), _p(lv), _p(_i32(up)), _p(_i32(ui)), _p(uv),
                                        dtype_code(lv.dtype), dtype, C.byref(opts) if opts is not None else None, C.byref(h)))
        return cls(ctx, h, n)

    @classmethod
    def split(cls, ctx, lp, li, lv, up, ui, uv, middle, lscale, rscale, permRow, permCol, dtype=F32, opts=None):
        n = len(lp) - 1
        lv = np.ascontiguousarray(lv)
        uv = np.ascontiguousarray(uv, dtype=lv.dtype)
        sd = np.float64
        middle, lscale, rscale = (np.ascontiguousarray(a, dtype=sd) for a in (middle, lscale, rscale))
        h = C.c_void_p()
        _chk(lib().bg_precond_create_split(ctx.h, n, _p(_i32(lp)), _p(_i32(li)), _p(lv), _p(_i32(up)), _p(_i32(ui)), _p(uv),
                                           dtype_code(lv.dtype), _p(middle), _p(lscale), _p(rscale), F64, _p(_i32(permRow)),
                                           _p(_i32(permCol)), dtype, C.byref(opts) if opts is not None else None, C.byref(h)))
        return cls(ctx, h, n)

    def apply(self, which, d_in, d_out):
        _chk(lib().bg_precond_apply(self.ctx.h, self.h, which, _p(d_in), _p(d_out)))

    def info(self) -> PrecondInfo:
        i = PrecondInfo()
        _chk(lib().bg_precond_get_info(self.h, C.byref(i)))
        return i

    def levels(self, is_lower):
        level = np.zeros(self.n, np.int32)
        _chk(lib().bg_precond_get_levels(self.h, int(is_lower), _p(level)))
        return level

    def close(self):
        if self.h:
            lib().bg_precond_destroy(self.h)
            self.h = C.c_void_p()


def gmres_opts(**kw) -> GmresOpts:
    o = GmresOpts()
    lib().bg_gmres_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Solver:
    def __init__(self, ctx, n, m=32, dtype=F64, basis_dtype=None):
        """basis_dtype=F32 with dtype=F64: the opt-in mixed-precision Krylov basis (bg_solver_create_ex)"""
        self.ctx, self.n, self.m, self.dtype = ctx, n, m, dtype
        self.h = C.c_void_p()
        if basis_dtype is None or basis_dtype == dtype:
            _chk(lib().bg_solver_create(ctx.h, n, m, dtype, C.byref(self.h)))
        else:
            _chk(lib().bg_solver_create_ex(ctx.h, n, m, dtype, basis_dtype, C.byref(self.h)))

    def solve(self, A, M, d_b, d_x, opts=None, host=False):
        """GMRES_GPU / GMRESilu_GPU through the C ABI.  d_b/d_x: device pointers (ints or torch
        data_ptr) unless host=True, in which case they are numpy arrays (x updated in place)."""
        o = opts if opts is not None else gmres_opts(m=self.m)
        r = GmresResult()
        f = lib().bg_gmres_solve_host if host else lib().bg_gmres_solve
        _chk(f(self.ctx.h, self.h, A.h, M.h, _p(d_b), _p(d_x), C.byref(o), C.byref(r)))
        return r

    def profile(self, A, M, d_b, d_x, opts=None):
        """Same solve with CUDA-event pairs around every SpMV / SpTRSV / orthogonalisation group."""
        o = opts if opts is not None else gmres_opts(m=self.m)
        r, st = GmresResult(), PhaseStats()
        _chk(lib().bg_solver_profile(self.ctx.h, self.h, A.h, M.h, _p(d_b), _p(d_x), C.byref(o), C.byref(r), C.byref(st)))
        return r, st

    def history(self):
        n = C.c_int(0)
        _chk(lib().bg_solver_get_history(self.h, None, 0, C.byref(n)))
        h = np.zeros(max(n.value, 1))
        _chk(lib().bg_solver_get_history(self.h, _p(h), n.value, C.byref(n)))
        return h[: n.value]

    def hessenberg(self, m=None):
        """m = restart length of the last solve (default: the workspace's)"""
        m = self.m if m is None else m
        H, s, cs, sn = np.zeros((m, m + 1)), np.zeros(m + 1), np.zeros(m + 1), np.zeros(m + 1)
        _chk(lib().bg_solver_get_hessenberg(self.h, _p(H), _p(s), _p(cs), _p(sn)))
        return H.T.copy(), s, cs, sn  # H[k, i]

    def bench_orth(self, j):
        _chk(lib().bg_solver_bench_orth(self.ctx.h, self.h, j))

    def mgs_fused(self) -> bool:
        return _chk(lib().bg_solver_mgs_fused(self.ctx.h, self.h)) == 1

    def close(self):
        if self.h:
            lib().bg_solver_destroy(self.h)
            self.h = C.c_void_p()


# ------------------------------------------------------------------ row-partitioned (multi-GPU) solve
class NetlistInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("n_inductors", C.c_int), ("n_vsources", C.c_int), ("n_isources", C.c_int),
                ("n_resistors", C.c_int), ("n_capacitors", C.c_int), ("size", C.c_int), ("size_mna", C.c_int),
                ("n_sources", C.c_int), ("n_ports", C.c_int), ("formulation", C.c_int), ("n_eliminated", C.c_int),
                ("n_pulse", C.c_int), ("dc", C.c_int), ("nnz_g", C.c_int64), ("nnz_c", C.c_int64), ("nnz_b", C.c_int64),
                ("n_pwl_points", C.c_int64), ("n_lines", C.c_int64), ("tstep", C.c_double), ("tstop", C.c_double)]


SRC_DC, SRC_PWL, SRC_PULSE = 0, 1, 2
NETLIST_MNA, NETLIST_REDUCED = 0, 1


class Netlist:
    """SPICE-subset power-grid netlist -> G, C, B (CSR triples), source tables, ports (bg_netlist_*): what
    Matrix / Transient take.  `source` is a path, or the netlist text itself when text=True."""

    def __init__(self, source, formulation=NETLIST_MNA, text=False, include_dir=None):
        self.h = C.c_void_p()
        if text:
            raw = source.encode() if isinstance(source, str) else bytes(source)
            _chk(lib().bg_netlist_parse(raw, C.c_int64(len(raw)), include_dir.encode() if include_dir else None, int(formulation), C.byref(self.h)))
        else:
            _chk(lib().bg_netlist_read(str(source).encode(), int(formulation), C.byref(self.h)))
        self.info = NetlistInfo()
        _chk(lib().bg_netlist_get_info(self.h, C.byref(self.info)))
        i = self.info
        self.G, self.C, self.B = (self._matrix(w, nz) for w, nz in (("G", i.nnz_g), ("C", i.nnz_c), ("B", i.nnz_b)))
        self.kind = np.zeros(i.n_sources, np.int32)
        self.pwl_ptr = np.zeros(i.n_sources + 1, np.int32)
        self.pwl_t, self.pwl_v = np.zeros(i.n_pwl_points), np.zeros(i.n_pwl_points)
        self.pulse_src = np.zeros(i.n_pulse, np.int32)
        self.pulse_time, self.pulse_val = np.zeros((i.n_pulse, 5)), np.zeros((i.n_pulse, 2))
        _chk(lib().bg_netlist_get_sources(self.h, _p(self.kind), _p(self.pwl_ptr), _p(self.pwl_t), _p(self.pwl_v), _p(self.pulse_src),
                                          _p(self.pulse_time), _p(self.pulse_val)))
        self.ports = np.zeros(i.n_ports, np.int32)
        _chk(lib().bg_netlist_get_ports(self.h, _p(self.ports)))
        self.kept = np.zeros(i.size, np.int32)
        self.elim_row, self.elim_src, self.elim_sign = np.zeros(i.n_eliminated, np.int32), np.zeros(i.n_eliminated, np.int32), np.zeros(i.n_eliminated)
        _chk(lib().bg_netlist_get_reduction(self.h, _p(self.kept), _p(self.elim_row), _p(self.elim_src), _p(self.elim_sign)))

    def _matrix(self, which, nnz):
        rp, ci, v = np.zeros(self.info.size + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        _chk(lib().bg_netlist_get_matrix(self.h, ord(which), _p(rp), _p(ci), _p(v)))
        return rp, ci, v

    def csc(self, which):
        """(colptr, row, val) with int64 indices: the cs_dl arrays the reference's parser_wrapper returns for 'G', 'C', 'B'"""
        i = self.info
        ncol = i.n_sources if which == "B" else i.size
        nnz = {"G": i.nnz_g, "C": i.nnz_c, "B": i.nnz_b}[which]
        cp, ri, x = np.zeros(ncol + 1, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz)
        _chk(lib().bg_netlist_get_matrix_csc(self.h, ord(which), _p(cp), _p(ri), _p(x)))
        return cp, ri, x

    @property
    def pwl(self):
        """per source (times, values): the `pwl` argument of Transient"""
        return [(self.pwl_t[a:b], self.pwl_v[a:b]) for a, b in zip(self.pwl_ptr[:-1], self.pwl_ptr[1:])]

    @property
    def transient_pwl(self):
        """`pwl` for Transient when the PULSE sources go through set_pulse: their tables shrink to the base value (the
        reference's six-points-per-period expansion is cut at tstop and need not be monotone in its last period)"""
        out = self.pwl
        for k, pv in zip(self.pulse_src, self.pulse_val):
            out[k] = (np.zeros(1), np.array([pv[0]]))
        return out

    def u(self, t):
        """source values at time t: piecewise-linear tables held at their end values outside; PULSE sources from their
        records like the device does (gen_PULSEut_kernel, src/kernels.cu:224-246) -- their tables are the reference's
        expansion, whose last breakpoint is cut to tstop and need not be monotone"""
        u = np.array([np.interp(t, tt, vv) if len(tt) else 0.0 for tt, vv in self.pwl])
        for k, (td, tr, tf, tw, tp), (vlo, vhi) in zip(self.pulse_src, self.pulse_time, self.pulse_val):
            mt = t - np.floor(t / tp) * tp
            if mt < td:
                u[k] = vlo
            elif mt < td + tr:
                u[k] = vlo + (mt - td) * (vhi - vlo) / tr
            elif mt < td + tr + tw:
                u[k] = vhi
            elif mt < td + tr + tw + tf:
                u[k] = vhi - (mt - td - tr - tw) * (vhi - vlo) / tf
            else:
                u[k] = vlo
        return u

    def node_row(self, name) -> int:
        """row of a node; -1 ground, -2 unknown, -3 eliminated by BG_NETLIST_REDUCED"""
        return int(lib().bg_netlist_node_row(self.h, str(name).encode()))

    def row_name(self, row) -> str:
        buf = C.create_string_buffer(256)
        _chk(lib().bg_netlist_row_name(self.h, int(row), buf, 256))
        return buf.value.decode()

    def expand(self, x, t=0.0):
        """solution of the active formulation -> MNA numbering (node voltages; the currents of eliminated sources stay NaN)"""
        full = np.full(self.info.size_mna, np.nan)
        full[self.kept] = x
        if len(self.elim_row):
            full[self.elim_row] = self.elim_sign * self.u(t)[self.elim_src]
        return full

    def close(self):
        if self.h:
            lib().bg_netlist_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close


class TransientStats(C.Structure):
    _fields_ = [("steps", C.c_int), ("iters_max", C.c_int), ("not_converged", C.c_int), ("iters_total", C.c_int64),
                ("launches", C.c_int64), ("last_resid", C.c_double), ("device_ms", C.c_double)]


class Transient:
    """Backward-Euler loop on the device: rhs = B u(t) + (C/h) x, warm-started GMRES, port sampling (bg_transient_*)."""

    def __init__(self, ctx, S, A, M, c_over_h, B, pwl, probes=()):
        """c_over_h, B: (rowptr, col, val) CSR triples (n x n and n x nsrc); pwl: list of (times, values) per source."""
        n = len(c_over_h[0]) - 1
        cv = np.ascontiguousarray(c_over_h[2], dtype=np.float64)
        bv = np.ascontiguousarray(B[2], dtype=np.float64)
        ptr = np.zeros(len(pwl) + 1, np.int32)
        ptr[1:] = np.cumsum([len(t) for t, _ in pwl])
        tt = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.float64) for t, _ in pwl]) if pwl else np.zeros(0))
        vv = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64) for _, v in pwl]) if pwl else np.zeros(0))
        pi = _i32(np.asarray(probes, dtype=np.int32))
        self.h, self.nprobe = C.c_void_p(), len(pi)
        _chk(lib().bg_transient_create(ctx.h, S.h, A.h, M.h, n, _p(_i32(c_over_h[0])), _p(_i32(c_over_h[1])), _p(cv), len(pwl),
                                       _p(_i32(B[0])), _p(_i32(B[1])), _p(bv), F64, _p(ptr), _p(tt), _p(vv), len(pi),
                                       _p(pi) if len(pi) else None, C.byref(self.h)))

    def set_pulse(self, src_index, pulse_time, pulse_val):
        """sources src_index[k] become PULSE sources: pulse_time[k] = (td, tr, tf, tw, tp), pulse_val[k] = (vlo, vhi)"""
        idx = np.ascontiguousarray(src_index, dtype=np.int32)
        pt = np.ascontiguousarray(pulse_time, dtype=np.float64).reshape(-1)
        pv = np.ascontiguousarray(pulse_val, dtype=np.float64).reshape(-1)
        assert len(pt) == 5 * len(idx) and len(pv) == 2 * len(idx)
        _chk(lib().bg_transient_set_pulse(self.h, len(idx), _p(idx), _p(pt), _p(pv)))

    def run(self, d_x, t0, h, nsteps, opts):
        st = TransientStats()
        probe = np.zeros((nsteps, self.nprobe)) if self.nprobe else None
        rc = lib().bg_transient_run(self.h, _p(d_x), C.c_double(t0), C.c_double(h), int(nsteps), C.byref(opts), C.byref(st),
                                    _p(probe) if probe is not None else None)
        if rc < 0:
            _chk(rc)
        return rc, st, probe

    def close(self):
        if self.h:
            lib().bg_transient_destroy(self.h)
            self.h = C.c_void_p()


class DistPlanInfo(C.Structure):
    _fields_ = [("nranks", C.c_int), ("rank", C.c_int), ("n_global", C.c_int64), ("n_local", C.c_int), ("n_halo", C.c_int),
                ("nnz_local", C.c_int64), ("n_neighbors", C.c_int), ("send_count", C.c_int64)]


class DistInfo(C.Structure):
    _fields_ = [("nranks", C.c_int), ("rank", C.c_int), ("n_global", C.c_int64), ("n_local", C.c_int), ("n_halo", C.c_int),
                ("n_neighbors", C.c_int), ("halo_messages", C.c_int64), ("allreduces", C.c_int64),
                ("halo_bytes_per_exchange", C.c_int64), ("precond", PrecondInfo), ("transport", C.c_int), ("p2p_available", C.c_int)]


TRANSPORT_NCCL, TRANSPORT_P2P = 0, 1


DIST_ID_BYTES = 128


class DistPlan:
    """One rank's share of a row-partitioned system (host only, no GPU needed)."""

    def __init__(self, rowptr, col, val, nranks, rank, part):
        n = len(rowptr) - 1
        val = np.ascontiguousarray(val)
        self.h = C.c_void_p()
        _chk(lib().bg_host_dist_plan(n, _p(_i32(rowptr)), _p(_i32(col)), _p(val), dtype_code(val.dtype), nranks, rank,
                                     _p(_i32(part)), C.byref(self.h)))
        self.info = DistPlanInfo()
        _chk(lib().bg_dist_plan_get_info(self.h, C.byref(self.info)))

    def arrays(self):
        i = self.info
        r = dict(l2g=np.zeros(i.n_local, np.int32), h2g=np.zeros(i.n_halo, np.int32), rowptr=np.zeros(i.n_local + 1, np.int32),
                 col=np.zeros(i.nnz_local, np.int32), val=np.zeros(i.nnz_local), nbr=np.zeros(i.n_neighbors, np.int32),
                 recv_ptr=np.zeros(i.n_neighbors + 1, np.int32), send_ptr=np.zeros(i.n_neighbors + 1, np.int32),
                 send_idx=np.zeros(i.send_count, np.int32))
        _chk(lib().bg_dist_plan_get(self.h, _p(r["l2g"]), _p(r["h2g"]), _p(r["rowptr"]), _p(r["col"]), _p(r["val"]), _p(r["nbr"]),
                                    _p(r["recv_ptr"]), _p(r["send_ptr"]), _p(r["send_idx"])))
        return r

    def close(self):
        if self.h:
            lib().bg_dist_plan_destroy(self.h)
            self.h = C.c_void_p()


def dist_unique_id() -> bytes:
    buf = C.create_string_buffer(DIST_ID_BYTES)
    _chk(lib().bg_dist_unique_id(buf))
    return buf.raw


class Dist:
    """Communicator + local matrix + per-rank block-Jacobi ILU + workspace (collective constructor)."""

    def __init__(self, ctx, plan, uid, dtype=F64, m=32, opts=None, ilut=False):
        self.ctx, self.h = ctx, C.c_void_p()
        idbuf = C.create_string_buffer(uid, DIST_ID_BYTES) if uid is not None else None
        _chk(lib().bg_dist_create(ctx.h, plan.h, idbuf, dtype, m, C.byref(opts) if opts is not None else None, int(ilut), C.byref(self.h)))
        self.n_local = plan.info.n_local

    def solve(self, d_b, d_x, opts, host=False):
        r = GmresResult()
        f = lib().bg_dist_gmres_solve_host if host else lib().bg_dist_gmres_solve
        _chk(f(self.h, _p(d_b), _p(d_x), C.byref(opts), C.byref(r)))
        return r

    def info(self) -> DistInfo:
        i = DistInfo()
        _chk(lib().bg_dist_get_info(self.h, C.byref(i)))
        return i

    def profile(self, d_b, d_x, opts):
        r, st = GmresResult(), PhaseStats()
        _chk(lib().bg_dist_profile(self.h, _p(d_b), _p(d_x), C.byref(opts), C.byref(r), C.byref(st)))
        return r, st

    def bench_comm(self, what, count=33, reps=200):
        """Device microseconds per halo exchange (what=0) or all-reduce of `count` doubles (what=1); collective."""
        us = C.c_double()
        _chk(lib().bg_dist_bench_comm(self.h, int(what), int(count), int(reps), C.byref(us)))
        return us.value

    def set_transport(self, transport):
        """TRANSPORT_P2P (peer-memory kernels over NVLink) or TRANSPORT_NCCL; same choice on every rank."""
        _chk(lib().bg_dist_set_transport(self.h, int(transport)))

    def close(self):
        if self.h:
            lib().bg_dist_destroy(self.h)
            self.h = C.c_void_p()
