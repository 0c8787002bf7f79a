#!/usr/bin/env python
"""bench.py -- time-to-solution of the GMRES(m)+ILU hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (CUDA library through the C ABI)
    python bench.py --impl reference --gpus 1 ...            # the reference's CPU path (ILU++ GMRES), same n

One "step" is one complete DC solve (relative residual 1e-8, GMRES(32)) of the synthetic 10 M-node 2-D power
grid BASELINE.json's north_star quotes the target on (pg2d 3162^2), with the factorisation done once outside the
timed region, exactly like the reference drivers reuse their preconditioner across solves
(src/mna_solve_gpu_gmres.cpp:518-531, 608-621).  `--workload ibmpg6` makes BASELINE configs[1] (1304^2) the headline.

Printed JSON (one line, rank 0): the contract keys plus
  roofline            dominant kernel (CUDA-event pairs around every launch group of a profiled solve,
                      bg_solver_profile) against MEASURED_PEAKS.json: `achieved`/`frac` on ALGORITHMIC bytes,
                      `moved_*` on the bytes the kernel actually moves (packed streams, bg_precond_info.stream_bytes)
  kernels             the same two fractions for every kernel group; iteration_roofline for the whole iteration
  reference_operator  N=1: the same solve with the reference's own operator (global ILU(0), level schedule) when the
                      headline runs block-Jacobi tiles, and vice versa (`block_jacobi`)
  reference_exact     N=1: level schedule + strict MGS (every choice the reference's) on the configs[1] grid, with
                      x_vs_oracle = ||x - x_oracle|| / ||x_oracle|| against the CPU oracle's solve of the same system
  configs1            BASELINE configs[1] as SURVEY 8 a20 defines it: ILU++ threshold factors (the reference's own
                      producer) through bg_precond_create_split = GMRESilu_GPU, ibmpg6-sized grid
  configs0            BASELINE configs[0]: ibmpg1-sized grid on the reference's compiled CPU GMRES()+MyILU0 (FP32,
                      oracle/_ref/libref_gmres.so) next to the same solve on the GPU
  cpu_baseline        the like-for-like CPU arm: the oracle's FP64 port of GMRES(32)+ILU(0) (src/gmres.cu:1904-2066),
                      one full solve of the configs[1]-sized grid, reported AT THAT SIZE (no extrapolation)
  e2e                 the same solve through the host-buffer entry point (bg_gmres_solve_host ≙
                      gmresInterfacePGfloat::GMRES_dev_PG): H2D of x and rhs, solve, D2H of x in the timed region.
Only the cpu_baseline / reference_exact check / configs0 / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
_JSON_FD = None  # N > 1: duplicate of the original stdout (descriptor 1 itself is pointed at stderr, see main)


def emit(line):
    """the ONE JSON line of this run, on the original stdout"""
    text = json.dumps(line) + "\n"
    if _JSON_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, text.encode())
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))

METRIC = "time_to_solve_relres_1e-8"
UNIT = "ms"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pg10m", choices=["pg10m", "ibmpg6"],
                    help="pg10m: the 10M-node power grid BASELINE.json's north_star quotes the target on (3162^2); "
                         "ibmpg6: BASELINE configs[1] (1304^2 = 1.7M nodes)")
    ap.add_argument("--nx", type=int, default=0, help="grid edge, overrides --workload; n = nx*nx")
    ap.add_argument("--quick", action="store_true", help="N=1: headline only (no secondary configs, no CPU legs)")
    ap.add_argument("--no-configs1", action="store_true", help="N=1: skip the ILU++-factor solve of BASELINE configs[1]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="N=1: also check the block-Jacobi solve of the configs[1] grid against the oracle given the same blocks")
    ap.add_argument("--precond", default="ilu0", choices=["ilu0", "ilut"])
    ap.add_argument("--schedule", default="block", choices=["block", "level"],
                    help="block: block-Jacobi ILU tiles (performance mode); level: the reference's global operator")
    ap.add_argument("--blocks", default="grid", choices=["grid", "metis"])
    ap.add_argument("--tile", type=int, default=32, help="block-Jacobi tile width in grid nodes (<= 32: one lane per tile column)")
    ap.add_argument("--tile-y", type=int, default=0, help="tile height (0 = same as --tile)")
    ap.add_argument("--ilut-tau", type=float, default=1e-2)
    ap.add_argument("--ilut-lfil", type=int, default=3)
    ap.add_argument("--factor-dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--orth", default="cgs", choices=["mgs", "cgs", "cgs2"])
    ap.add_argument("--basis", default="f64", choices=["f64", "f32"], help="storage of the Krylov basis (f32: opt-in mixed precision, N=1, CGS)")
    ap.add_argument("--eta", type=float, default=0.1, help="CGS re-orthogonalisation threshold (DGKS): second pass when ||w'|| < eta ||w||")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="N>1: peer-memory kernels over NVLink (default) or the NCCL baseline")
    a = ap.parse_args()
    if a.tile_y <= 0:
        a.tile_y = a.tile
    if a.nx <= 0:
        a.nx = 3162 if a.workload == "pg10m" else 1304
    return a


def workload_name(nx):
    if nx == 3162:
        return "pg2d 3162x3162 = synthetic 10M-node power grid (the workload BASELINE.json's north_star states the target on)"
    if nx == 1304:
        return "pg2d 1304x1304 (ibmpg6-sized, BASELINE configs[1])"
    if nx == 174:
        return "pg2d 174x174 (ibmpg1-sized, BASELINE configs[0])"
    return f"pg2d {nx}x{nx}"


def config_of(nx, n, nnz, tol):
    """Identical for both arms (the driver compares them): what is solved, not how."""
    ws = (33 * 8 + 12 * 5 + 32) * n / 1e6  # basis + matrix + factor stream, FP64, MB
    l2 = ("no flush between timed solves: one solve streams its working set (matrix, factors, 33 basis vectors of n values: "
          "%.0f MB in FP64) many times over, far beyond the 126 MB L2" % ws) if ws > 2 * 126 else (
          "working set %.0f MB: L2-resident, a parity-size run, not a bandwidth measurement" % ws)
    return {"workload": f"{workload_name(nx)} DC solve, GMRES, rel. residual {tol:g}", "n": int(n), "nnz": int(nnz), "tol": tol, "l2_policy": l2}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.monotonic()] + [c.strip() for c in line.split(",")])

    def mark(self):
        return time.monotonic()

    def stop(self, t0=None, t1=None):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        rows = [r[1:] for r in self.rows if len(r) >= 10]
        window = [r[1:] for r in self.rows if len(r) >= 10 and t0 is not None and t0 <= r[0] <= t1 + 0.05]
        scope = "timed region"
        if len(window) < 2:  # region shorter than two sampling periods: use everything sampled under load
            window, scope = rows, "warm-up + timed region"
        sm = [float(r[1]) for r in window if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in window if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({nm for r in window for nm, c in zip(names, r[5:9]) if c.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "scope": scope}


_WL = {}


def workload(nx):
    from b200gmres import synth

    if nx not in _WL:
        _WL.clear()  # one large grid at a time
        _WL[nx] = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 2)
    return _WL[nx]


# --------------------------------------------------------------------------- CPU legs (the only users of oracle/)
def cpu_like_for_like(nx, tol):
    """The like-for-like CPU arm: the oracle's FP64 port of the reference's host GMRES(32) + MyILU0::HostPrecond
    (src/gmres.cu:1904-2066, src/preconditioner.cu:205-251), one full DC solve, single thread like the reference.
    Reported at the size actually run.  Returns (dict, x_oracle)."""
    from oracle import pyoracle as O

    O.build()
    rp, ci, v, b = workload(nx)
    n = len(rp) - 1
    t0 = time.perf_counter()
    f = O.ilu0(rp, ci, v, "f64")
    setup_s = time.perf_counter() - t0
    pc = O.Precond(O.PC_ILU0, n, "f64", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
    t0 = time.perf_counter()
    flag, x, iters, res, _ = O.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, tol, pc)
    ms = 1e3 * (time.perf_counter() - t0)
    return ({"value": ms, "unit": UNIT, "cores": 1, "kind": "port", "n": int(n), "extrapolated": False,
             "sample": f"one full DC solve of the {nx}x{nx} grid (n={n}): oracle FP64 port of GMRES(32)+ILU(0) "
                       f"(gmres.cu:1904-2066 + MyILU0::HostPrecond), {iters} iterations, resid {res:.2e}; the value is for THIS n, not scaled",
             "setup_s": setup_s, "iters": int(iters), "resid": float(res), "host_cores_online": os.cpu_count()}, x, (int(flag), int(iters), float(res)))


def cpu_ilupp(nx, tol):
    """The reference's `-gmres -ilu` CPU path (src/mna_solve_gpu_gmres.cpp:1896-2276): ILU++ ILUCDP (threshold 1.2)
    + iluplusplus::gmres, double, single-threaded, from the vendored ILU++ compiled unchanged into
    oracle/_ref/libilupp.so.  One timed solve; factorisation outside the timed region (reported as setup_s)."""
    import scipy.sparse as sp

    from oracle import pyoracle as O

    if O.ilupp() is None:
        return None
    rp, ci, v, b = workload(nx)
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc()
    A.sort_indices()
    t0 = time.perf_counter()
    fac = O.ilupp_factor(n, A.indptr, A.indices, A.data, threshold=1.2, mem_factor=1.0)
    setup_s = time.perf_counter() - t0
    # restart = 100 is the reference driver's own setting for this path (mna_solve_gpu_gmres.cpp:2063); with restart 32
    # the locally patched ILU++ gmres never recomputes r at a restart and stagnates.  ILU++ ANDs a
    # relative and an absolute test (both -log10 exponents): the absolute one is switched off so the stopping rule
    # is the metric's relative residual.
    t0 = time.perf_counter()
    x, iters, rt, at = O.ilupp_gmres(fac, b, restart=100, max_iter=1000, rel_tol=-np.log10(tol), abs_tol=-30.0)
    ms = 1e3 * (time.perf_counter() - t0)
    O.ilupp_free(fac)
    true_rel = float(np.linalg.norm(b - _spmv_host(rp, ci, v, x)) / np.linalg.norm(b))
    return {"value": ms, "unit": UNIT, "cores": 1, "kind": "reference", "n": int(n), "extrapolated": False,
            "sample": f"one full DC solve of the {nx}x{nx} grid (n={n}), not scaled: ILU++ ILUCDP(threshold 1.2) + "
                      f"iluplusplus::gmres(restart 100) double, {iters} iterations, reached 1e-{rt:.2f}",
            "setup_s": setup_s, "iters": int(iters), "true_rel_residual": true_rel, "host_cores_online": os.cpu_count()}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, on the SAME
    grid as our arm.  A solve of this size takes about a minute on one core (the reference is single-threaded), so ONE
    solve is timed whatever --steps / --warmup say; the line reports steps = 1, warmup = 0."""
    if rank != 0:
        return
    rp, ci, v, b = workload(args.nx)
    n = len(rp) - 1
    cb = cpu_ilupp(args.nx, args.tol)
    if cb is None:  # ILU++ not built: the oracle port of GMRES(32)+ILU(0)
        cb = cpu_like_for_like(args.nx, args.tol)[0]
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": cb["value"], "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.nx, n, len(ci), args.tol),
            "impl_config": {"algorithm": cb["sample"], "requested_steps": args.steps, "requested_warmup": args.warmup,
                            "note": "one timed sample: a single-threaded CPU solve of this grid takes of the order of a minute"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    if world > 1:
        # rank 0 prints ONE JSON line on stdout.  NCCL logs to the process's stdout, so file descriptor 1 is pointed at
        # stderr for everything but that line (kept on a duplicate of the original descriptor); the log runs at INFO so
        # that the communicator's rank count and the transports it picked are visible to whoever runs this
        global _JSON_FD
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):  # (this image exports NCCL_DEBUG=VERSION)
            os.environ["NCCL_DEBUG"] = os.environ.get("BG_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        run_dist(args, rank, world, local)
        return

    line = run_single(args, local, args.nx, "ilu0" if args.precond == "ilu0" else "ilut", args.schedule, args.orth, args.steps, headline=True,
                      keep_x=not args.quick, basis=args.basis)
    x_head = line.pop("_x", None)
    if not args.quick and args.basis == "f64" and args.orth != "mgs":
        # opt-in mixed-precision Krylov basis (FP32 columns under FP64 vectors) on the headline system: never the headline,
        # reported next to it with its distance from the all-FP64 run
        mb = run_single(args, local, args.nx, "ilu0" if args.precond == "ilu0" else "ilut", args.schedule, args.orth, min(args.steps, 5),
                        headline=False, keep_x=True, basis="f32")
        sm = _summary(mb, extra=("iteration_roofline",))
        sm["x_vs_fp64_basis"] = float(np.linalg.norm(mb["_x"] - x_head) / np.linalg.norm(x_head))
        sm["true_rel_residual_fp64_basis"] = line["true_rel_residual"]
        sm["iters_fp64_basis"] = line["iters"]
        line["fp32_basis_opt_in"] = sm
        # the reference's own arithmetic and stopping rule: FP32 everywhere, tolerance 1e-6 (src/gmres.h:55,
        # docs/tvlsi_3d_thermal.pdf p.578), same grid, same schedule
        rp32 = run_single(args, local, args.nx, "ilu0", args.schedule, args.orth, min(args.steps, 5), headline=False, keep_x=True,
                          vec="f32", tol=1e-6)
        s32 = _summary(rp32, extra=("iteration_roofline",))
        s32["x_vs_fp64_headline"] = float(np.linalg.norm(rp32["_x"] - x_head) / np.linalg.norm(x_head))
        line["reference_precision_fp32_tol1e-6"] = s32
    if not args.quick:
        other = "level" if args.schedule == "block" else "block"
        sec = run_single(args, local, args.nx, "ilu0", other, args.orth, min(args.steps, 5), headline=False)
        line["reference_operator" if other == "level" else "block_jacobi"] = _summary(sec, extra=("iteration_roofline",))
        # configs[1]-sized grid: the fully reference-exact configuration, checked against the oracle's solve
        x_oracle = None
        if not args.no_cpu_baseline:
            line["cpu_baseline"], x_oracle, ometa = cpu_like_for_like(1304, args.tol)
        ex = run_single(args, local, 1304, "ilu0", "level", "mgs", min(args.steps, 5), headline=False, keep_x=True)
        rex = _summary(ex)
        if x_oracle is not None:
            rex["x_vs_oracle"] = float(np.linalg.norm(ex["_x"] - x_oracle) / np.linalg.norm(x_oracle))
            rex["oracle"] = {"flag": ometa[0], "iters": ometa[1], "resid": ometa[2]}
            rex["iters_equal_oracle"] = bool(ex["iters"] == ometa[1])
        line["reference_exact"] = rex
        if args.nx != 1304:
            bj = run_single(args, local, 1304, "ilu0", "block", args.orth, min(args.steps, 10), headline=False, keep_x=True)
            s1 = _summary(bj)
            if args.verify:
                s1["x_vs_oracle_same_blocks"] = verify_block(1304, args, bj["_x"])
            line["ibmpg6_block_jacobi"] = s1
        if not args.no_configs1:
            line["configs1"] = run_split_ilupp(args, local, 1304, min(args.steps, 10))
        line["configs0"] = run_configs0(args, local)
    emit(line)


def _summary(sec, extra=()):
    keys = ("value", "unit", "iters", "ms_per_iteration", "true_rel_residual", "kernels", "e2e", "setup_s", "config", "impl_config") + tuple(extra)
    return {k: sec[k] for k in keys if k in sec}


def verify_block(nx, args, x):
    """oracle GMRES with the block-Jacobi ILU(0) of the SAME tiles"""
    import b200gmres as bg
    from oracle import pyoracle as O

    rp, ci, v, b = workload(nx)
    n = len(rp) - 1
    nb, blk = bg.host_block_partition(bg.PART_GRID2D, rp, ci, 0, [nx, nx, 1, args.tile, args.tile_y, 1])
    f = bg.host_ilu0(rp, ci, v, blk)
    pc = O.Precond(O.PC_ILU0, n, "f64", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
    flag, xo, it, res, _ = O.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, args.tol, pc)
    return {"x_vs_oracle": float(np.linalg.norm(x - xo) / np.linalg.norm(xo)), "oracle_iters": int(it)}


def _time_solves(torch, ctx, stream, fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    e0.record(stream)
    for _ in range(k):
        r = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), r, ctx.launches - l0


def run_single(args, local, nx, precond, schedule, orth_name, steps, headline, keep_x=False, basis="f64", vec="f64", tol=None):
    """vec = "f32": vectors, matrix and factors in FP32 (the reference's own arithmetic, src/gmres.h:55)"""
    import torch

    import b200gmres as bg

    rp, ci, v, b = workload(nx)
    n = len(rp) - 1
    tol = args.tol if tol is None else tol
    vdt, npdt, tdt, s8 = (bg.F64, np.float64, torch.float64, 8) if vec == "f64" else (bg.F32, np.float32, torch.float32, 4)
    ctx = bg.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    t_matrix = time.perf_counter()
    A = bg.Matrix(ctx, rp, ci, v, dtype=vdt)
    t_matrix = time.perf_counter() - t_matrix
    fdt = bg.F32 if (args.factor_dtype == "f32" or vec == "f32") else bg.F64
    t_setup = time.perf_counter()
    if schedule == "block":
        o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=args.tile * args.tile_y, factor_dtype=fdt,
                            block_method=bg.PART_GRID2D if args.blocks == "grid" else bg.PART_METIS,
                            dims=[nx, nx, 1, args.tile, args.tile_y, 1], ilut_tau=args.ilut_tau, ilut_lfil=args.ilut_lfil)
    else:
        o = bg.precond_opts(trsv_schedule=bg.TRSV_LEVEL, factor_dtype=fdt, ilut_tau=args.ilut_tau, ilut_lfil=args.ilut_lfil)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=vdt, opts=o, ilut=(precond == "ilut"))
    t_setup = time.perf_counter() - t_setup
    S = bg.Solver(ctx, n, 32, dtype=vdt, basis_dtype=bg.F32 if (basis == "f32" and vec == "f64") else None)
    orth = {"mgs": bg.ORTH_MGS, "cgs": bg.ORTH_CGS_ADAPTIVE, "cgs2": bg.ORTH_CGS2}[orth_name]
    opts = bg.gmres_opts(m=32, tol=tol, max_iter=10000, orth=orth, reorth_eta=args.eta)

    bv = np.ascontiguousarray(b, dtype=npdt)
    db = torch.from_numpy(bv).cuda()
    dx = torch.zeros(n, dtype=tdt, device="cuda")
    hb = torch.from_numpy(bv).pin_memory()
    hx = torch.zeros(n, dtype=tdt).pin_memory()
    torch.cuda.synchronize()

    def step_device():
        with torch.cuda.stream(stream):
            dx.zero_()  # DC solve starts from x = 0 (mna_solve_gpu_gmres.cpp:525-527)
        return S.solve(A, M, db.data_ptr(), dx.data_ptr(), opts)

    def step_host():
        hx.zero_()
        return S.solve(A, M, hb.numpy(), hx.numpy(), opts, host=True)

    sampler = ClockSampler(local)
    if headline:
        sampler.start()
        time.sleep(1.0)  # nvidia-smi needs a moment before its first sample
    t_first = time.perf_counter()
    step_device()  # the first solve also builds the block-major copy of A (on the device)
    ctx.sync()
    t_first = time.perf_counter() - t_first
    for _ in range((max(args.warmup, 3) if headline else 2) - 1):
        step_device()
    tc0 = sampler.mark()
    ms, r, launches = _time_solves(torch, ctx, stream, step_device, steps)
    tc1 = sampler.mark()
    clocks = sampler.stop(tc0, tc1) if headline else None
    for _ in range(2 if headline else 1):
        step_host()
    ms_e2e, r_e2e, _ = _time_solves(torch, ctx, stream, step_host, steps)

    # profiled solve: CUDA-event pairs around every launch group, on the launching stream
    with torch.cuda.stream(stream):
        dx.zero_()
    rprof, st = S.profile(A, M, db.data_ptr(), dx.data_ptr(), opts)
    x = dx.cpu().numpy().astype(np.float64)
    true_rel = float(np.linalg.norm(b - _spmv_host(rp, ci, v, x)) / np.linalg.norm(b))

    peak, which = peaks()
    info, pinfo = A.info(), M.info()
    # bytes one preconditioner application really moves: the packed factor streams + the vectors it reads and writes
    # (block tiles: r in, z out; level strips: r in, tmp out, tmp in, z out; sync-free slices add the 16-byte packets)
    if schedule == "block":
        trsv_moved_per = pinfo.stream_bytes + 2 * s8 * n
        trsv_kernel = "trsv_block_fused (L and U solve of every diagonal block, one launch)"
    elif pinfo.lattice:
        trsv_moved_per = pinfo.stream_bytes + 4 * s8 * n
        trsv_kernel = "trsv_lattice_global (global 5-point ILU(0): wavefront strips, L launch + U launch)"
    else:
        trsv_moved_per = pinfo.stream_bytes + 4 * s8 * n + 2 * 16 * n
        trsv_kernel = "trsv_syncfree (global factors: level slices, flag-in-data packets, L launch + U launch)"
    napply = st.trsv_calls
    # SpMV moves its sliced-ELL copy (values + 32-bit columns, padded) + x gathers (L2-served) + y
    spmv_moved_per = (info.sell_padded_nnz * (s8 + 4) if info.format == bg.SPMV_SELL else info.nnz * (s8 + 4) + 4 * (n + 1)) + 2 * s8 * n
    kernels = {
        ("spmv_sell_w<%s,5> (sliced ELL, fixed width, all loads in flight)" % ("double" if vec == "f64" else "float") if info.format == bg.SPMV_SELL else "spmv_csr_vector"):
            (st.spmv_ms, st.spmv_bytes, st.spmv_calls, spmv_moved_per * st.spmv_calls),
        trsv_kernel: (st.trsv_ms, st.trsv_bytes, st.trsv_calls, trsv_moved_per * napply),
    }
    if orth_name == "mgs":
        it = max(rprof.iters_done, 1)
        if S.mgs_fused():
            # one cooperative launch per inner iteration: every basis column read once, w read once, v_{i+1} written once
            moved = sum(((t % 32 + 1) + 2) * s8 * n for t in range(it))
            kernels["mgs_kernel (strict MGS, i+1 dependent dot/axpy pairs in one launch; w in registers, packet all-to-all sums)"] = (
                st.orth_ms, st.orth_bytes, st.orth_calls, moved)
        else:
            # i+1 (dot, axpy) launch pairs per inner iteration; each pair reads v_k and w twice and writes w
            moved = sum((5 * (t % 32 + 1) + 3) * s8 * n for t in range(it))
            kernels["dots_kernel+update_kernel (strict MGS, i+1 pairs)"] = (st.orth_ms, st.orth_bytes, st.orth_calls, moved)
    elif st.dots_calls == 0:
        # fused step: one cooperative launch per inner iteration (projections + Givens epilogue + v_j)
        ob = st.dots_bytes + st.update_bytes
        if basis == "f32":  # (2j) 4-byte basis reads, w twice, the new vector once as f64 and once as f32
            it = max(rprof.iters_done, 1)
            ob = sum((2 * (t % 32 + 1) * 4 + 2 * s8 + s8 + 4) * n for t in range(it))
        kernels["orth_kernel (fused: V^T w, w.w, Givens, v_j = (w - V h)/||.||; TMA-staged tiles reused from shared memory / L2)"] = (
            st.orth_ms, st.dots_bytes + st.update_bytes, st.orth_calls, ob)
    else:
        kernels["dots_kernel (V^T w and w.w, TMA-staged tiles)"] = (st.dots_ms, st.dots_bytes, st.dots_calls, st.dots_bytes)
        kernels["update_kernel (w - V h, normalised, TMA-staged tiles)"] = (st.update_ms, st.update_bytes, st.update_calls, st.update_bytes)
    tot_ms = st.spmv_ms + st.trsv_ms + st.orth_ms + st.other_ms
    dom = max(kernels, key=lambda k: kernels[k][0])
    dms, dbytes, dcalls, dmoved = kernels[dom]
    ach = dbytes / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
    traffic = ncu_traffic(dom, "orth_kernel<double,float> (opt-in FP32 basis)" if (basis == "f32" and vec == "f64" and dom.startswith("orth_kernel")) else None)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "frac_of_nominal_7700": ach / 7700.0,
                "moved_GBps": dmoved / (dms * 1e-3) / 1e9 if dms > 0 else 0.0, "frac_moved": (dmoved / (dms * 1e-3) / 1e9 / peak) if dms > 0 else 0.0,
                "traffic": traffic.get("bytes") if traffic else None, "traffic_launch": traffic,
                "peak_source": which, "bytes_per_launch": dbytes / max(dcalls, 1), "us_per_launch": 1e3 * dms / max(dcalls, 1),
                "launches_timed": int(dcalls), "share_of_step": dms / tot_ms if tot_ms else None,
                "how": "CUDA-event pairs on the library stream around every launch of the profiled solve (bg_solver_profile); "
                       "achieved/frac = SURVEY 8(d) algorithmic bytes, moved_* = bytes the kernel actually moves"}
    per_kernel = {}
    for k, p in kernels.items():
        g_alg = p[1] / (p[0] * 1e-3) / 1e9 if p[0] > 0 else 0.0
        g_mov = p[3] / (p[0] * 1e-3) / 1e9 if p[0] > 0 else 0.0
        per_kernel[k] = {"ms": p[0], "launches": int(p[2]), "us_per_launch": 1e3 * p[0] / max(p[2], 1),
                         "GBps_algorithmic": g_alg, "frac_algorithmic": g_alg / peak, "GBps_moved": g_mov, "frac_moved": g_mov / peak,
                         "share": p[0] / tot_ms if tot_ms else None}
    step_ms = ms / steps
    alg_bytes = st.spmv_bytes + st.trsv_bytes + st.orth_bytes + st.other_bytes
    moved_bytes = sum(p[3] for p in kernels.values()) + st.other_bytes
    itp = max(rprof.iters_done, 1)
    line = {
        "metric": METRIC, "value": step_ms, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3) if headline else 2,
        "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": vec, "data": "synthetic",
        "config": config_of(nx, n, info.nnz, tol),
        "impl_config": {"restart": 32, "precond": precond, "trsv_schedule": schedule, "lattice_layout": bool(pinfo.lattice),
                        "operator": ("block-Jacobi ILU tiles (changes the preconditioner: iteration drift reported through iters)" if schedule == "block"
                                     else "the reference's global factors, getLevel dependencies"),
                        "blocks": pinfo.nblocks, "block_rows": pinfo.max_block_rows, "levels": [pinfo.levelsL, pinfo.levelsU],
                        "factor_dtype": "f32" if fdt == bg.F32 else "f64", "basis_dtype": basis if vec == "f64" else "f32", "orth": orth_name, "reorth_eta": args.eta,
                        "spmv_format": "sell" if info.format == bg.SPMV_SELL else "csr",
                        "l2_policy": "working set (Krylov basis 33 x n x %d B = %.0f MB) exceeds the 126 MB L2" % (s8, 33 * n * s8 / 1e6)},
        "iters": int(r.iters), "resid": float(r.resid), "true_rel_residual": true_rel, "reorths": int(r.reorths),
        "ms_per_iteration": step_ms / max(r.iters_done, 1),
        "iteration_roofline": {"bytes_per_iteration_moved": moved_bytes / itp, "moved_GBps": moved_bytes / (step_ms * 1e-3) / 1e9,
                               "frac": moved_bytes / (step_ms * 1e-3) / 1e9 / peak,
                               "bytes_per_iteration_survey": alg_bytes / itp,
                               "frac_survey_count": alg_bytes / (step_ms * 1e-3) / 1e9 / peak,
                               "frac_of_nominal_7700": moved_bytes / (step_ms * 1e-3) / 1e9 / 7700.0,
                               "note": "frac = bytes actually moved per solve (packed streams, vectors) / time / measured COPY bandwidth; "
                                       "frac_survey_count charges SURVEY 8(d)'s general-CSR count (12 B per factor entry + indices), "
                                       "which the index-free layouts do not move"},
        "kernels": per_kernel, "setup_s": t_setup,
        "setup_breakdown_s": {"matrix_upload_and_spmv_storage": t_matrix, "preconditioner": t_setup, "first_solve": t_first,
                              "note": "host wall clock of bg_matrix_create_csr, bg_precond_create_*, and the first bg_gmres_solve "
                                      "(which includes the one-time block-major reordering); all outside the timed region"},
        "roofline": roofline,
        "e2e": {"value": ms_e2e / steps, "unit": UNIT, "h2d_bytes_per_step": 2 * n * s8, "d2h_bytes_per_step": n * s8},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if keep_x:
        line["_x"] = x
    S.close(); M.close(); A.close(); ctx.close()
    del db, dx, hb, hx
    torch.cuda.empty_cache()
    return line


def run_split_ilupp(args, local, nx, steps):
    """BASELINE configs[1] as SURVEY 8 a20 defines it: factors from the reference's own ILU++ (threshold 1.2,
    src/mna_solve_gpu_gmres.cpp:316-344) handed over like setPrecondPG does, GMRESilu_GPU algebra (split
    preconditioning), GMRES(32).  ILU++ runs on the CPU (it is the reference's factor producer), outside the timed region."""
    import scipy.sparse as sp
    import torch

    import b200gmres as bg
    from oracle import pyoracle as O

    if O.ilupp() is None:
        return {"unavailable": "oracle/_ref/libilupp.so is not built"}
    rp, ci, v, b = workload(nx)
    n = len(rp) - 1
    Ac = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc()
    Ac.sort_indices()
    t0 = time.perf_counter()
    fac = O.ilupp_factor(n, Ac.indptr, Ac.indices, Ac.data, threshold=1.2, mem_factor=1.0)
    t_ilupp = time.perf_counter() - t0
    keys = ("lp", "li", "lv", "up", "ui", "uv", "middle", "lscale", "rscale", "permRow", "permCol")
    ctx = bg.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
    t0 = time.perf_counter()
    M = bg.Precond.split(ctx, *[fac[k] for k in keys], dtype=bg.F64)
    t_dev = time.perf_counter() - t0
    nnzL, nnzU = len(fac["lv"]), len(fac["uv"])
    O.ilupp_free(fac)
    S = bg.Solver(ctx, n, 32, dtype=bg.F64)
    out = {}
    for orth_name, orth in (("mgs", bg.ORTH_MGS), ("cgs", bg.ORTH_CGS_ADAPTIVE)):
        opts = bg.gmres_opts(m=32, tol=args.tol, max_iter=10000, orth=orth, reorth_eta=args.eta)
        db = torch.from_numpy(b).cuda()
        dx = torch.zeros(n, dtype=torch.float64, device="cuda")

        def step():
            with torch.cuda.stream(stream):
                dx.zero_()
            return S.solve(A, M, db.data_ptr(), dx.data_ptr(), opts)

        step()
        ms, r, _ = _time_solves(torch, ctx, stream, step, steps)
        with torch.cuda.stream(stream):
            dx.zero_()
        rprof, st = S.profile(A, M, db.data_ptr(), dx.data_ptr(), opts)
        x = dx.cpu().numpy()
        pinfo = M.info()
        out[orth_name] = {"value": ms / steps, "unit": UNIT, "iters": int(r.iters), "resid": float(r.resid),
                          "true_rel_residual": float(np.linalg.norm(b - _spmv_host(rp, ci, v, x)) / np.linalg.norm(b)),
                          "ms_per_iteration": ms / steps / max(r.iters_done, 1),
                          "phases_ms": {"spmv": st.spmv_ms, "trsv_L_and_U": st.trsv_ms, "orth": st.orth_ms, "restart_terms": st.other_ms},
                          "trsv_us_per_solve": 1e3 * st.trsv_ms / max(st.trsv_calls, 1), "trsv_solves": int(st.trsv_calls)}
    pinfo = M.info()
    out["config"] = config_of(nx, n, len(ci), args.tol)
    out["impl_config"] = {"restart": 32, "precond": "ILU++ multilevel ILUCDP, threshold 1.2 (the reference's producer), split: M_L = L, M_R = D U",
                          "algebra": "GMRESilu_GPU (src/gmres.cu:2254-2446)", "nnzL": nnzL, "nnzU": nnzU, "levels": [pinfo.levelsL, pinfo.levelsU],
                          "kernel": "trsv_syncfree"}
    out["setup_s"] = {"ilupp_cpu": t_ilupp, "upload_and_schedule": t_dev}
    S.close(); M.close(); A.close(); ctx.close()
    torch.cuda.empty_cache()
    return out


def run_configs0(args, local):
    """BASELINE configs[0]: ibmpg1-sized grid (174^2 = 30 276 nodes), GMRES(32)+ILU(0) on the reference's CPU path --
    the reference's own compiled host GMRES() + MyILU0::HostPrecond (src/gmres.cu:1904-2066, FP32, tolerance 1e-6 =
    its driver default, src/defs.h:12) from oracle/_ref/libref_gmres.so -- next to the same FP32 solve on the GPU."""
    import torch

    import b200gmres as bg
    from b200gmres import synth
    from oracle import pyoracle as O

    rp, ci, v, b = synth.pg2d(174, 174, seed=synth.SEED_BASE + 1)
    n = len(rp) - 1
    out = {"config": config_of(174, n, len(ci), 1e-6)}
    O.build()
    f32 = O.ilu0(rp, ci, v, "f32")
    times = []
    kind = "reference"
    for _ in range(5):
        t0 = time.perf_counter()
        if O.ref_gmres_lib() is not None:
            flag, x, it, res = O.ref_host_gmres_ilu0(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-6, f32)[:4]
        else:
            kind = "port"
            pc = O.Precond(O.PC_ILU0, n, "f32", **{k: f32[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
            flag, x, it, res, _ = O.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-6, pc)
        times.append(time.perf_counter() - t0)
    out["cpu"] = {"value": 1e3 * float(np.median(times)), "unit": UNIT, "kind": kind, "cores": 1, "iters": int(it), "resid": float(res),
                  "sample": "5 full solves, median; the reference's compiled GMRES() + MyILU0::HostPrecond, FP32, tol 1e-6"}
    ctx = bg.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F32)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F32)
    S = bg.Solver(ctx, n, 32, dtype=bg.F32)
    opts = bg.gmres_opts(m=32, tol=1e-6, max_iter=10000, orth=bg.ORTH_MGS)
    hb = torch.from_numpy(b.astype(np.float32)).pin_memory()
    hx = torch.zeros(n, dtype=torch.float32).pin_memory()

    def step_host():
        hx.zero_()
        return S.solve(A, M, hb.numpy(), hx.numpy(), opts, host=True)

    step_host()
    ms, r, _ = _time_solves(torch, ctx, stream, step_host, 10)
    xg = hx.numpy().astype(np.float64)
    out["gpu"] = {"value": ms / 10, "unit": UNIT, "iters": int(r.iters), "resid": float(r.resid), "dtype": "f32",
                  "what": "bg_gmres_solve_host (H2D, solve, D2H), level schedule (reference operator), strict MGS",
                  "x_vs_cpu_reference": float(np.linalg.norm(xg - x) / np.linalg.norm(x))}
    S.close(); M.close(); A.close(); ctx.close()
    return out


def run_dist(args, rank, world, local):
    """N > 1: the same DC solve, row-partitioned into `world` horizontal strips of the grid (contiguous
    row blocks, src/partition.cpp's initial assignment), one process per GPU: halo exchange per SpMV, one
    all-reduce per orthogonalisation step, per-rank block-Jacobi ILU (dist.cu)."""
    import torch
    import torch.distributed as dist

    import b200gmres as bg

    dev = torch.device("cuda", local)
    rp, ci, v, b = workload(args.nx)
    n, nx = len(rp) - 1, args.nx
    # strips of whole grid rows, heights as equal as possible
    ys = [(r * nx) // world for r in range(world + 1)]
    part = np.repeat(np.arange(world, dtype=np.int32), np.diff(ys) * nx)
    t_setup = time.perf_counter()
    plan = bg.DistPlan(rp, ci, v, world, rank, part)
    t_plan = time.perf_counter() - t_setup
    ids = [bg.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = bg.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    fdt = bg.F32 if args.factor_dtype == "f32" else bg.F64
    ny_local = ys[rank + 1] - ys[rank]
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=args.tile * args.tile_y, factor_dtype=fdt, block_method=bg.PART_GRID2D,
                        dims=[nx, ny_local, 1, args.tile, args.tile_y, 1], ilut_tau=args.ilut_tau, ilut_lfil=args.ilut_lfil)
    D = bg.Dist(ctx, plan, ids[0], dtype=bg.F64, m=32, opts=o, ilut=(args.precond == "ilut"))
    t_setup = time.perf_counter() - t_setup
    if args.transport == "nccl" or not D.info().p2p_available:
        D.set_transport(bg.TRANSPORT_NCCL)
    a = plan.arrays()
    nl = plan.info.n_local
    orth = {"mgs": bg.ORTH_MGS, "cgs": bg.ORTH_CGS_ADAPTIVE, "cgs2": bg.ORTH_CGS2}[args.orth]
    opts = bg.gmres_opts(m=32, tol=args.tol, max_iter=10000, orth=orth, reorth_eta=args.eta)
    bl = np.ascontiguousarray(b[a["l2g"]])
    db = torch.from_numpy(bl).to(dev)
    dx = torch.zeros(nl, dtype=torch.float64, device=dev)
    hb = torch.from_numpy(bl).pin_memory()
    hx = torch.zeros(nl, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()

    def step_device():
        with torch.cuda.stream(stream):
            dx.zero_()
        return D.solve(db.data_ptr(), dx.data_ptr(), opts)

    def step_host():
        hx.zero_()
        return D.solve(hb.numpy(), hx.numpy(), opts, host=True)

    def timed(fn, k):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record(stream)
        for _ in range(k):
            r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # device time, max over ranks
        return float(t.item()), r, ctx.launches - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(1.0)
    for _ in range(max(args.warmup, 3)):
        step_device()
    tc0 = sampler.mark()
    ms, r, launches = timed(step_device, args.steps)
    tc1 = sampler.mark()
    clocks = sampler.stop(tc0, tc1) if rank == 0 else None
    for _ in range(2):
        step_host()
    ms_e2e, _, _ = timed(step_host, args.steps)
    # per-phase device time of one more collective solve (event pairs on every rank's stream; rank 0 reports)
    with torch.cuda.stream(stream):
        dx.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    rprof, ph = D.profile(db.data_ptr(), dx.data_ptr(), opts)
    itp = max(rprof.iters_done, 1)
    phases = {"us_per_iteration": {"halo": 1e3 * ph.halo_ms / itp, "halo+spmv": 1e3 * ph.spmv_ms / itp, "trsv": 1e3 * ph.trsv_ms / itp, "dots(+allreduce+givens)": 1e3 * ph.dots_ms / itp,
                                   "update": 1e3 * ph.update_ms / itp, "orth_total(fused kernel when dots=0)": 1e3 * ph.orth_ms / itp, "restart_terms": 1e3 * ph.other_ms / itp},
              "note": "rank 0, CUDA-event pairs around every launch group of a separate profiled solve (adds event gaps)"}
    dist.barrier()
    comm_us = {"halo_exchange": D.bench_comm(0), "allreduce_33_doubles_in_dots_kernel": D.bench_comm(1, 33)}
    # true residual of the assembled solution (gathered on rank 0)
    step_device()
    torch.cuda.synchronize()
    sizes = [int((part == q).sum()) for q in range(world)]
    pieces = []
    for q in range(world):
        t = dx.clone() if q == rank else torch.zeros(sizes[q], dtype=torch.float64, device=dev)
        dist.broadcast(t, src=q)
        pieces.append(t)
    info = D.info()
    if rank == 0:
        x = np.concatenate([p.cpu().numpy() for p in pieces])
        true_rel = float(np.linalg.norm(b - _spmv_host(rp, ci, v, x)) / np.linalg.norm(b))
        peak, which = peaks()
        step_ms = ms / args.steps
        s8 = 8
        nnz = len(ci)
        it = max(r.iters_done, 1)
        # per inner iteration, all ranks: SURVEY 8(d) count and the bytes the kernels actually move (index-free tile
        # streams: rank 0's stream x world)
        orth_b = sum((2 * (t % 32 + 1) + 5) * s8 * n for t in range(it))
        spmv_b = ((s8 + 4) * nnz + 4 * (n + 1) + 2 * s8 * n) * it
        trsv_b = ((s8 + 4) * (info.precond.nnzL + info.precond.nnzU) * world + 8 * (n + 1) + 4 * s8 * n) * it
        trsv_m = (info.precond.stream_bytes * world + 2 * s8 * n) * it
        all_b, all_m = orth_b + spmv_b + trsv_b, orth_b + spmv_b + trsv_m
        line = {
            "metric": METRIC, "value": step_ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(nx, n, nnz, args.tol),
            "impl_config": {"restart": 32, "precond": args.precond, "trsv_schedule": "block",
                            "partition": f"{world} horizontal strips (contiguous rows), per-rank block-Jacobi ILU, tiles {args.tile}x{args.tile_y}",
                            "orth": args.orth, "reorth_eta": args.eta,
                            "l2_policy": "per-rank Krylov basis 33 x n/N x 8 B = %.0f MB" % (33 * n * 8 / 1e6 / world)},
            "iters": int(r.iters), "resid": float(r.resid), "true_rel_residual": true_rel, "reorths": int(r.reorths),
            "ms_per_iteration": step_ms / it,
            "iteration_roofline": {"bytes_per_iteration_moved": all_m / it, "moved_GBps": all_m / (step_ms * 1e-3) / 1e9,
                                   "frac": all_m / (step_ms * 1e-3) / 1e9 / (peak * world), "peak": peak * world,
                                   "bytes_per_iteration_survey": all_b / it, "frac_survey_count": all_b / (step_ms * 1e-3) / 1e9 / (peak * world),
                                   "note": "all ranks; peak = N x measured HBM copy bandwidth; frac on bytes actually moved"},
            "comm": {"halo_bytes_per_exchange_rank0": int(info.halo_bytes_per_exchange), "neighbors_rank0": int(info.n_neighbors),
                     "allreduce_doubles_max": 34,
                     "backend": ("peer-memory kernels over NVLink (CUDA IPC windows): all-reduce fused into dots_kernel, halo = one gather+peer-store kernel"
                                 if info.transport == bg.TRANSPORT_P2P else "NCCL send/recv + allreduce on the solver stream")},
            "phases_rank0": phases, "comm_us_per_op_rank0": comm_us,
            "roofline": None, "setup_s": t_setup, "setup_plan_s": t_plan,
            "e2e": {"value": ms_e2e / args.steps, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": n * 8},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(line)
    D.close(); plan.close(); ctx.close()
    dist.destroy_process_group()


def ncu_traffic(kernel, prefer=None):
    """DRAM bytes of one captured launch of the dominant kernel (ncu --set full, read offline into
    profiles/ncu_traffic.json by tools/ncu_summary.py --traffic); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    table = json.load(open(p))
    if prefer and prefer in table:
        return table[prefer]
    for key, val in table.items():
        if kernel.startswith(key) or key in kernel:
            return val
    return None


def _spmv_host(rp, ci, v, x):
    import scipy.sparse as sp

    return sp.csr_matrix((v, ci, rp), shape=(len(rp) - 1, len(rp) - 1)) @ x


if __name__ == "__main__":
    main()
