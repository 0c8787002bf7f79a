"""Generate the committed golden fixtures under tests/golden/ (run in the build container).

    python tests/golden/make_golden.py

Everything is produced by the CPU oracle (oracle/oracle.cpp, a restatement of the reference's
host twins) on seeded synthetic inputs; the ILU++ factors of the split-preconditioner cases come
from the reference's own vendored ILU++ 1.1.1 compiled unchanged into oracle/_ref/libilupp.so
(needs /root/reference, which the GPU box does not have -- hence committed fixtures).
The reference ships no golden vectors of its own for this path.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
from b200gmres import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "pg2d_24x20": lambda: synth.pg2d(24, 20, seed=synth.SEED_BASE + 11),
    "th3d_8x7x6": lambda: synth.th3d(8, 7, 6, seed=synth.SEED_BASE + 13),
    "rand_300": lambda: synth.random_sparse(300, 6, seed=synth.SEED_BASE + 17),
}


def ilu0_case(name, gen):
    rp, ci, v, b = gen()
    n = len(rp) - 1
    out = dict(rowptr=rp, col=ci, val=v, b=b)
    for mode, tol in (("f32", 1e-6), ("f64", 1e-8)):
        f = O.ilu0(rp, ci, v, mode)
        for k in ("lp", "li", "lv", "up", "ui", "uv", "level", "levelPtr", "nodes"):
            out[f"{mode}_{k}"] = f[k]
        out[f"{mode}_levL"] = O.get_level(1, f["lp"], f["li"], f["lv"], mode)
        out[f"{mode}_levU"] = O.get_level(0, f["up"], f["ui"], f["uv"], mode)
        pc = O.Precond(O.PC_ILU0, n, mode, lp=f["lp"], li=f["li"], lv=f["lv"], up=f["up"], ui=f["ui"], uv=f["uv"])
        flag, x, it, res, hist = O.gmres(rp, ci, v, b, np.zeros(n), 32, 2000, tol, pc, hist_cap=2000)
        out[f"{mode}_x"], out[f"{mode}_hist"] = x, hist
        out[f"{mode}_meta"] = np.array([flag, it, res, tol])
        # one preconditioner application on the rhs (forward + backward solve)
        out[f"{mode}_Minv_b"] = pc.apply(0, b)
        # restart exercised: small m
        flag, x, it, res, hist = O.gmres(rp, ci, v, b, np.ones(n), 5, 2000, tol, pc, hist_cap=2000)
        out[f"{mode}_m5_x"] = x
        out[f"{mode}_m5_meta"] = np.array([flag, it, res, tol])
    A = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc()
    A.sort_indices()
    for npart in (2, 4):
        ps, npt, pinv, q = O.partition(A.indptr, A.indices, npart, n)
        out[f"part{npart}_size"], out[f"part{npart}_node"], out[f"part{npart}_pinv"], out[f"part{npart}_q"] = ps, npt, pinv, q
    # unpreconditioned and Jacobi runs (MyNONE / MyDIAG)
    pc = O.Precond(O.PC_NONE, n, "f64")
    flag, x, it, res, hist = O.gmres(rp, ci, v, b, np.zeros(n), 32, 400, 1e-8, pc, hist_cap=400)
    out["none_x"], out["none_meta"] = x, np.array([flag, it, res, 1e-8])
    dinv = 1.0 / sp.csr_matrix((v, ci, rp), shape=(n, n)).diagonal()
    pc = O.Precond(O.PC_DIAG, n, "f64", dinv=dinv)
    flag, x, it, res, hist = O.gmres(rp, ci, v, b, np.zeros(n), 32, 400, 1e-8, pc, hist_cap=400)
    out["diag_x"], out["diag_meta"] = x, np.array([flag, it, res, 1e-8])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, n, "iters f32/f64", int(out["f32_meta"][1]), int(out["f64_meta"][1]))


def split_case(name, gen):
    """ILU++ (threshold 1.2, mna_solve_gpu_gmres.cpp:316-329) factors + GMRESilu on them."""
    rp, ci, v, b = gen()
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc()
    A.sort_indices()
    fac = O.ilupp_factor(n, A.indptr, A.indices, A.data, threshold=1.2, mem_factor=1.0)
    out = dict(rowptr=rp, col=ci, val=v, b=b, fill=fac["fill"])
    keys = ("lp", "li", "lv", "up", "ui", "uv", "middle", "lscale", "rscale", "permRow", "permCol")
    for k in keys:
        out[k] = fac[k]
    kw = {k: fac[k] for k in keys}
    for mode, tol in (("f32f64", 1e-7), ("f64", 1e-8)):
        pc = O.Precond(O.PC_SPLIT, n, mode, **kw)
        flag, x, it, res, hist = O.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, tol, pc, hist_cap=1000)
        out[f"{mode}_x"], out[f"{mode}_hist"], out[f"{mode}_meta"] = x, hist, np.array([flag, it, res, tol])
        # warm start from a perturbed solution (transient pattern, gmres_interface_pg.cu:118)
        x0 = x * (1 + 1e-3 * np.cos(np.arange(n)))
        flag, xw, it, res, hist = O.gmres(rp, ci, v, b, x0, 32, 10000, tol, pc, hist_cap=1000)
        out[f"{mode}_warm_x0"], out[f"{mode}_warm_x"], out[f"{mode}_warm_meta"] = x0, xw, np.array([flag, it, res, tol])
        for which, nm in ((1, "left"), (2, "right"), (3, "start")):
            out[f"{mode}_{nm}_b"] = pc.apply(which, b)
    x, it, rt, at = O.ilupp_gmres(fac, b)
    out["ilupp_x"], out["ilupp_iters"] = x, it
    O.ilupp_free(fac)
    np.savez_compressed(os.path.join(OUT, name + "_split.npz"), **out)
    print(name, "split fill", out["fill"], "iters", int(out["f32f64_meta"][1]), int(out["f64_meta"][1]), "ilupp", it)


if __name__ == "__main__":
    O.build()
    for name, gen in CASES.items():
        ilu0_case(name, gen)
    for name in ("pg2d_24x20", "th3d_8x7x6"):
        split_case(name, CASES[name])
