"""ctypes door onto oracle/liboracle.so and oracle/_ref/*.so -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PC_NONE, PC_DIAG, PC_ILU0, PC_SPLIT = 0, 1, 2, 3

_DT = {"f32": np.float32, "f64": np.float64, "f32f64": np.float32}
_FT = {"f32": np.float32, "f64": np.float64, "f32f64": np.float64}


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) or (
        os.path.isdir("/root/reference/src") and not os.path.exists(os.path.join(_HERE, "_ref", "libilupp.so"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-j4"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None
_ilupp = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_HERE, "liboracle.so"))
    return _lib


def ref():
    """Reference's own CPU translation units (oracle/_ref/libref.so); None if not built."""
    global _ref
    p = os.path.join(_HERE, "_ref", "libref.so")
    if _ref is None and os.path.exists(p):
        _ref = C.CDLL(p)
    return _ref


def ilupp():
    global _ilupp
    p = os.path.join(_HERE, "_ref", "libilupp.so")
    if _ilupp is None and os.path.exists(p):
        _ilupp = C.CDLL(p)
        _ilupp.ilupp_create.restype = C.c_void_p
    return _ilupp


_ref_gmres = None


def ref_gmres_lib():
    """The reference's own gmres.cu / preconditioner.cu / leftILU.cu compiled unchanged (oracle/_ref/libref_gmres.so)."""
    global _ref_gmres
    p = os.path.join(_HERE, "_ref", "libref_gmres.so")
    if _ref_gmres is None and os.path.exists(p):
        _ref_gmres = C.CDLL(p)
    return _ref_gmres


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _PcStruct(C.Structure):
    _fields_ = [("kind", C.c_int), ("n", C.c_int), ("dinv", C.c_void_p), ("lv", C.c_void_p), ("uv", C.c_void_p),
                ("lp", C.c_void_p), ("li", C.c_void_p), ("up", C.c_void_p), ("ui", C.c_void_p),
                ("middle", C.c_void_p), ("lscale", C.c_void_p), ("rscale", C.c_void_p),
                ("permRow", C.c_void_p), ("permCol", C.c_void_p)]


class Precond:
    """Flat description of a reference Preconditioner object for the oracle."""

    def __init__(self, kind, n, mode="f64", **kw):
        self.kind, self.n, self.mode = kind, n, mode
        R, F = _DT[mode], _FT[mode]
        self._keep = {}
        s = _PcStruct()
        s.kind, s.n = kind, n
        for name, dt in (("dinv", R), ("lv", F), ("uv", F), ("middle", R), ("lscale", R), ("rscale", R)):
            if kw.get(name) is not None:
                self._keep[name] = np.ascontiguousarray(kw[name], dtype=dt)
                setattr(s, name, self._keep[name].ctypes.data)
        for name in ("lp", "li", "up", "ui", "permRow", "permCol"):
            if kw.get(name) is not None:
                self._keep[name] = _i32(kw[name])
                setattr(s, name, self._keep[name].ctypes.data)
        self.struct = s

    def apply(self, which, vec):
        R = _DT[self.mode]
        vin = np.ascontiguousarray(vec, dtype=R)
        out = np.zeros_like(vin)
        getattr(lib(), "orc_precond_apply_" + self.mode)(C.byref(self.struct), C.c_int(which), _p(vin), _p(out))
        return out


def spmv(rowptr, col, val, x, mode="f64"):
    R = _DT[mode]
    val = np.ascontiguousarray(val, dtype=R); x = np.ascontiguousarray(x, dtype=R)
    out = np.zeros(len(rowptr) - 1, dtype=R)
    getattr(lib(), "orc_spmv_" + ("f32" if R == np.float32 else "f64"))(
        _p(out), _p(val), _p(_i32(rowptr)), _p(_i32(col)), _p(x), C.c_int(len(rowptr) - 1))
    return out


def generate_level(rowptr, col, val, mode="f32"):
    """leftILU.cu generateLevel + bucketing -> (level, levelPtr, nodes)."""
    R = _DT[mode]; n = len(rowptr) - 1
    val = np.ascontiguousarray(val, dtype=R)
    level = np.zeros(n, dtype=np.int32)
    getattr(lib(), "orc_generate_level_" + mode)(C.c_int(n), _p(val), _p(_i32(rowptr)), _p(_i32(col)), _p(level))
    return (level,) + bucket_levels(level)


def bucket_levels(level):
    level = _i32(level); n = len(level)
    nl = int(level.max()) + 1 if n else 0
    lp = np.zeros(nl + 1, dtype=np.int32); nodes = np.zeros(n, dtype=np.int32)
    lib().orc_bucket_levels(C.c_int(n), _p(level), _p(lp), _p(nodes))
    return lp, nodes


def get_level(is_l, rowptr, col, val, mode="f32"):
    R = _DT[mode]; n = len(rowptr) - 1
    val = np.ascontiguousarray(val, dtype=R)
    level = np.zeros(n, dtype=np.int32)
    getattr(lib(), "orc_get_level_" + mode)(C.c_int(int(is_l)), _p(val), _p(_i32(rowptr)), _p(_i32(col)), _p(level), C.c_int(n))
    return level


def ilu0(rowptr, col, val, mode="f32"):
    """leftILU CPU semantics.  Returns dict(lp,li,lv,up,ui,uv,level,levelPtr,nodes,fval,fcol)."""
    R = _DT[mode]; n = len(rowptr) - 1; nnz = int(rowptr[-1])
    rowptr, col = _i32(rowptr), _i32(col); val = np.ascontiguousarray(val, dtype=R)
    f = getattr(lib(), "orc_ilu0_" + mode)
    lp = np.zeros(n + 1, dtype=np.int32); up = np.zeros(n + 1, dtype=np.int32)
    nl = C.c_int(0)
    f(C.c_int(n), _p(val), _p(rowptr), _p(col), _p(lp), None, None, _p(up), None, None, None, C.byref(nl), None, None, None, None)
    lv = np.zeros(lp[-1], dtype=R); li = np.zeros(lp[-1], dtype=np.int32)
    uv = np.zeros(up[-1], dtype=R); ui = np.zeros(up[-1], dtype=np.int32)
    level = np.zeros(n, dtype=np.int32); levelPtr = np.zeros(nl.value + 1, dtype=np.int32)
    nodes = np.zeros(n, dtype=np.int32); fval = np.zeros(nnz, dtype=R); fcol = np.zeros(nnz, dtype=np.int32)
    f(C.c_int(n), _p(val), _p(rowptr), _p(col), _p(lp), _p(lv), _p(li), _p(up), _p(uv), _p(ui), _p(level),
      C.byref(nl), _p(levelPtr), _p(nodes), _p(fval), _p(fcol))
    return dict(lp=lp, li=li, lv=lv, up=up, ui=ui, uv=uv, level=level, levelPtr=levelPtr, nodes=nodes,
                fval=fval, fcol=fcol)


def gmres(rowptr, col, val, b, x0, m, max_iter, tol, pc: Precond, hist_cap=0):
    """Reference host GMRES / GMRESilu.  Returns (flag, x, iters, resid, hist)."""
    mode = pc.mode; R = _DT[mode]; n = len(rowptr) - 1
    val = np.ascontiguousarray(val, dtype=R); b = np.ascontiguousarray(b, dtype=R)
    x = np.array(x0, dtype=R, copy=True)
    it = C.c_int(max_iter)
    t = (C.c_float if R == np.float32 else C.c_double)(tol)
    hist = np.zeros(max(hist_cap, 1), dtype=R); hl = C.c_int(0)
    fn = getattr(lib(), "orc_gmres_" + mode); fn.restype = C.c_int
    flag = fn(_p(val), _p(_i32(rowptr)), _p(_i32(col)), _p(x), _p(b), C.c_int(n), C.c_int(m), C.byref(it),
              C.byref(t), C.byref(pc.struct), _p(hist), C.c_int(hist_cap), C.byref(hl))
    return flag, x, it.value, t.value, hist[: hl.value].copy()


def partition(Ap, Ai, npart, n_nodes):
    m = len(Ap) - 1
    Ap = np.ascontiguousarray(Ap, dtype=np.int64); Ai = np.ascontiguousarray(Ai, dtype=np.int64)
    ps = np.zeros(npart + 1, dtype=np.int64); npt = np.zeros(m, dtype=np.int64)
    pinv = np.zeros(m, dtype=np.int64); q = np.zeros(m, dtype=np.int64)
    lib().orc_partition(C.c_int64(m), _p(Ap), _p(Ai), C.c_int(npart), C.c_int64(n_nodes), _p(ps), _p(npt), _p(pinv), _p(q))
    return ps, npt, pinv, q


def ilupp_factor(n, Ap, Ai, Ax, threshold=1.2, mem_factor=1.0):
    """Run the reference's vendored ILU++ (oracle/_ref/libilupp.so) like mna_solve_gpu_gmres.cpp:316-405."""
    L = ilupp()
    if L is None:
        raise RuntimeError("oracle/_ref/libilupp.so not built")
    Ap = _i32(Ap); Ai = _i32(Ai); Ax = np.ascontiguousarray(Ax, dtype=np.float64)
    h = C.c_void_p(L.ilupp_create(C.c_int(n), _p(Ap), _p(Ai), _p(Ax), C.c_double(threshold), C.c_double(mem_factor)))
    nl, nu, fill = C.c_int(0), C.c_int(0), C.c_double(0)
    L.ilupp_sizes(h, C.byref(nl), C.byref(nu), C.byref(fill))
    r = dict(lp=np.zeros(n + 1, np.int32), li=np.zeros(nl.value, np.int32), lv=np.zeros(nl.value),
             up=np.zeros(n + 1, np.int32), ui=np.zeros(nu.value, np.int32), uv=np.zeros(nu.value),
             middle=np.zeros(n), permRow=np.zeros(n, np.int32), permCol=np.zeros(n, np.int32),
             lscale=np.zeros(n), rscale=np.zeros(n))
    L.ilupp_get(h, *[_p(r[k]) for k in ("lp", "li", "lv", "up", "ui", "uv", "middle", "permRow", "permCol", "lscale", "rscale")])
    r["fill"] = fill.value
    r["_handle"] = h
    return r


def ilupp_gmres(fac, b, restart=100, max_iter=6000, rel_tol=8.0, abs_tol=8.0):
    L = ilupp(); n = len(b)
    b = np.ascontiguousarray(b, dtype=np.float64); x = np.zeros(n)
    rt, at = C.c_double(rel_tol), C.c_double(abs_tol)
    L.ilupp_gmres.restype = C.c_int
    it = L.ilupp_gmres(fac["_handle"], _p(b), _p(x), C.c_int(restart), C.c_int(max_iter), C.byref(rt), C.byref(at))
    return x, it, rt.value, at.value


def ilupp_free(fac):
    ilupp().ilupp_destroy(fac.pop("_handle"))


# --------------------------------------------------------------------------- the reference's compiled host solver
def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ref_host_gmres_ilu0(rowptr, col, val, b, x0, m, max_iter, tol, f):
    """Reference GMRES() (src/gmres.cu:1904-2066) + MyILU0::HostPrecond, REAL = float, factors f (dict of ilu0())."""
    L = ref_gmres_lib(); n = len(rowptr) - 1
    x = np.array(x0, dtype=np.float32, copy=True)
    it, t = C.c_int(max_iter), C.c_float(tol)
    lv, uv = _f32(f["lv"]).copy(), _f32(f["uv"]).copy()
    lp, li, up, ui = (_i32(f[k]).copy() for k in ("lp", "li", "up", "ui"))
    L.ref_gmres_ilu0.restype = C.c_int
    flag = L.ref_gmres_ilu0(C.c_int(n), _p(_f32(val)), _p(_i32(rowptr)), _p(_i32(col)), _p(lv), _p(lp), _p(li), _p(uv), _p(up), _p(ui),
                            _p(x), _p(_f32(b)), C.c_int(m), C.byref(it), C.byref(t))
    return flag, x, it.value, t.value


def ref_host_gmres_none(rowptr, col, val, b, x0, m, max_iter, tol):
    L = ref_gmres_lib(); n = len(rowptr) - 1
    x = np.array(x0, dtype=np.float32, copy=True)
    it, t = C.c_int(max_iter), C.c_float(tol)
    L.ref_gmres_none.restype = C.c_int
    flag = L.ref_gmres_none(C.c_int(n), _p(_f32(val)), _p(_i32(rowptr)), _p(_i32(col)), _p(x), _p(_f32(b)), C.c_int(m), C.byref(it), C.byref(t))
    return flag, x, it.value, t.value


def ref_host_ilu0_apply(f, vec):
    L = ref_gmres_lib(); n = len(f["lp"]) - 1
    out = np.zeros(n, dtype=np.float32)
    lv, uv = _f32(f["lv"]).copy(), _f32(f["uv"]).copy()
    lp, li, up, ui = (_i32(f[k]).copy() for k in ("lp", "li", "up", "ui"))
    L.ref_ilu0_hostprecond(C.c_int(n), _p(lv), _p(lp), _p(li), _p(uv), _p(up), _p(ui), _p(_f32(vec)), _p(out))
    return out


def _split_args(g):
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64).copy()
    return [f64(g["lv"]), _i32(g["lp"]).copy(), _i32(g["li"]).copy(), f64(g["uv"]), _i32(g["up"]).copy(), _i32(g["ui"]).copy(),
            _f32(g["middle"]).copy(), _i32(g["permRow"]).copy(), _i32(g["permCol"]).copy(), _f32(g["lscale"]).copy(), _f32(g["rscale"]).copy()]


def ref_host_gmresilu_split(rowptr, col, val, b, x0, m, max_iter, tol, g):
    """Reference GMRESilu() (src/gmres.cu:2069-2252) + MyILUPPfloat host twins on ILU++ factors g (double), float vectors."""
    L = ref_gmres_lib(); n = len(rowptr) - 1
    x = np.array(x0, dtype=np.float32, copy=True)
    it, t = C.c_int(max_iter), C.c_float(tol)
    a = _split_args(g)
    L.ref_gmresilu_split.restype = C.c_int
    flag = L.ref_gmresilu_split(C.c_int(n), _p(_f32(val)), _p(_i32(rowptr)), _p(_i32(col)), *[_p(v) for v in a], _p(x), _p(_f32(b)),
                                C.c_int(m), C.byref(it), C.byref(t))
    return flag, x, it.value, t.value


def ref_host_split_apply(which, g, vec):
    """which: 1 left, 2 right, 3 starting_value, 4 rhs (MyILUPPfloat::HostPrecond_*)."""
    L = ref_gmres_lib(); n = len(g["lp"]) - 1
    out = np.zeros(n, dtype=np.float32)
    a = _split_args(g)
    L.ref_split_hostprecond(C.c_int(which), C.c_int(n), *[_p(v) for v in a], _p(_f32(vec)), _p(out))
    return out


def ref_get_level(is_l, rowptr, col, val):
    L = ref_gmres_lib(); n = len(rowptr) - 1
    level = np.zeros(n, dtype=np.int32)
    L.ref_getLevel(C.c_int(int(is_l)), _p(_f32(val)), _p(_i32(rowptr)), _p(_i32(col)), _p(level), C.c_int(n))
    return level


def ref_generate_level(rowptr, col, val):
    L = ref_gmres_lib(); n = len(rowptr) - 1
    node = np.zeros(n, dtype=np.int32); level = np.zeros(n, dtype=np.int32)
    L.ref_generateLevel(C.c_int(n), _p(_f32(val)), _p(_i32(rowptr)), _p(_i32(col)), _p(node), _p(level))
    return node, level


def ref_left_ilu(rowptr, col, val):
    """Reference leftILU() (src/leftILU.cu:27-336).  Needs a CUDA device (it cudaMallocs even when every level runs on the CPU)."""
    L = ref_gmres_lib(); n = len(rowptr) - 1; nnz = int(rowptr[-1])
    rp, ci, v = _i32(rowptr).copy(), _i32(col).copy(), _f32(val).copy()
    lp, up = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    li, ui = np.zeros(nnz + n, np.int32), np.zeros(nnz + n, np.int32)
    lv, uv = np.zeros(nnz + n, np.float32), np.zeros(nnz + n, np.float32)
    L.ref_leftILU(C.c_int(n), _p(v), _p(rp), _p(ci), _p(lv), _p(lp), _p(li), _p(uv), _p(up), _p(ui))
    return dict(lp=lp, li=li[: lp[-1]], lv=lv[: lp[-1]], up=up, ui=ui[: up[-1]], uv=uv[: up[-1]])


# --------------------------------------------------------------------------- the reference's file readers
_ref_io = None


def ref_io_lib():
    global _ref_io
    p = os.path.join(_HERE, "_ref", "libref_io.so")
    if _ref_io is None and os.path.exists(p):
        _ref_io = C.CDLL(p)
    return _ref_io


def ref_read_sparse_matrix(path):
    """readSparseMatrix(format 0) + genCSRFormat (src_thermal/SpMV_gen.cpp:93-201) -> (rowptr, col, val float32)."""
    L = ref_io_lib()
    nr, nc, nnz = C.c_int(), C.c_int(), C.c_int()
    L.ref_readSparseMatrix(str(path).encode(), C.byref(nr), C.byref(nc), C.byref(nnz), None, None, None)
    val = np.zeros(nnz.value, np.float32); rp = np.zeros(nr.value + 1, np.int32); ci = np.zeros(nnz.value, np.int32)
    L.ref_readSparseMatrix(str(path).encode(), C.byref(nr), C.byref(nc), C.byref(nnz), _p(val), _p(rp), _p(ci))
    return rp, ci, val, nc.value


def ref_read_input_vector(path, n):
    y = np.zeros(n, np.float32)
    ref_io_lib().ref_readInputVector(_p(y), str(path).encode(), C.c_int(n))
    return y


def ref_write_output_vector(path, x):
    x = _f32(x).copy()
    ref_io_lib().ref_writeOutputVector(_p(x), str(path).encode(), C.c_int(len(x)))


def ref_write_csr(path, rowptr, col, val, n_cols):
    rp, ci, v = _i32(rowptr).copy(), _i32(col).copy(), _f32(val).copy()
    ref().ref_writeCSR(C.c_int(len(rp) - 1), C.c_int(n_cols), C.c_int(len(ci)), _p(rp), _p(ci), _p(v), str(path).encode())
