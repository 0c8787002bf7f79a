#!/usr/bin/env python
"""bench.py -- time-to-solution of the GMRES(m)+ILU hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (CUDA library through the C ABI)
    python bench.py --impl reference --gpus 1 --steps 5 ...  # the reference's CPU path (ILU++ GMRES)

One "step" is one complete DC solve (relative residual 1e-8, GMRES(32)) of the synthetic
ibmpg6-sized 2-D power grid (BASELINE.json configs[1]; n = 1304^2 = 1 700 416 unknowns) with the
factorisation done once outside the timed region, exactly like the reference drivers reuse their
preconditioner across solves (src/mna_solve_gpu_gmres.cpp:518-531, 608-621).

Printed JSON (one line, rank 0): the contract keys plus
  roofline      dominant kernel group (from CUDA-event pairs around every launch group of a
                profiled solve, bg_solver_profile) against MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU implementation (oracle/_ref/libilupp.so = vendored ILU++
                GMRES, kind "reference") or the oracle port, timed on this box's host cores
  e2e           the same solve through the host-buffer entry point (bg_gmres_solve_host ≙
                gmresInterfacePGfloat::GMRES_dev_PG): H2D of x and rhs, solve, D2H of x in the
                timed region.
Only the cpu_baseline / --impl reference legs touch oracle/.
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
    ap.add_argument("--no-configs1", action="store_true", help="N=1: skip the secondary BASELINE configs[1] solve (ibmpg6-sized, ILUT)")
    ap.add_argument("--precond", default="ilu0", choices=["ilu0", "ilut"])
    ap.add_argument("--schedule", default="block", choices=["block", "level"])
    ap.add_argument("--blocks", default="grid", choices=["grid", "metis"])
    ap.add_argument("--tile", type=int, default=32, help="block-Jacobi tile width in grid nodes (<= 32: one lane per tile column)")
    ap.add_argument("--tile-y", type=int, default=0, help="tile height (0 = same as --tile)")
    ap.add_argument("--ilut-tau", type=float, default=1e-2)
    ap.add_argument("--ilut-lfil", type=int, default=3)
    ap.add_argument("--factor-dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--orth", default="cgs", choices=["mgs", "cgs", "cgs2"])
    ap.add_argument("--eta", type=float, default=0.1, help="CGS re-orthogonalisation threshold (DGKS): second pass when ||w'|| < eta ||w||")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="N>1: peer-memory kernels over NVLink (default) or the NCCL baseline")
    ap.add_argument("--cpu-nx", type=int, default=652, help="grid edge of the bounded CPU sample (0 = same as --nx); 652^2 = n/4")
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
    return f"pg2d {nx}x{nx}"


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


def workload(nx):
    from b200gmres import synth

    rp, ci, v, b = synth.pg2d(nx, nx, seed=synth.SEED_BASE + 2)
    return rp, ci, v, b


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference(nx_full, nx_cpu, steps, warmup, tol):
    """The reference's CPU implementation of the path, timed on the host cores of this box.

    kind "reference": iluplusplus::gmres with the ILUCDP preconditioner, the `-gmres -ilu` driver
    path (src/mna_solve_gpu_gmres.cpp:1896-2276), from the vendored ILU++ compiled unchanged into
    oracle/_ref/libilupp.so.  Single-threaded, as in the reference.  Falls back to the oracle port
    of GMRES + ILU(0) (kind "port") when that library is absent."""
    import scipy.sparse as sp

    from oracle import pyoracle as O

    nx = nx_cpu or nx_full
    rp, ci, v, b = workload(nx)
    n = len(rp) - 1
    scale = (nx_full * nx_full) / float(n)
    times, iters, kind, setup_s = [], None, "port", 0.0
    if O.ilupp() is not None:
        kind = "reference"
        A = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc()
        A.sort_indices()
        t0 = time.perf_counter()
        fac = O.ilupp_factor(n, A.indptr, A.indices, A.data, threshold=1.2, mem_factor=1.0)
        setup_s = time.perf_counter() - t0
        # restart = 100 is the reference driver's own setting for this path (mna_solve_gpu_gmres.cpp:2063);
        # with restart 32 the locally patched ILU++ gmres never recomputes r at a restart
        # and stagnates.  ILU++ ANDs a relative and an absolute test (both -log10 exponents): the
        # absolute one is switched off so the stopping rule is the metric's relative residual.
        ex = -np.log10(tol)
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            x, iters, rt, at = O.ilupp_gmres(fac, b, restart=100, max_iter=1000, rel_tol=ex, abs_tol=-30.0)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
        O.ilupp_free(fac)
        what = f"ILU++ ILUCDP(threshold 1.2) + iluplusplus::gmres(restart 100) double, {iters} iterations, reached 1e-{rt:.2f}"
    else:
        O.build()
        t0 = time.perf_counter()
        f = O.ilu0(rp, ci, v, "f64")
        setup_s = time.perf_counter() - t0
        pc = O.Precond(O.PC_ILU0, n, "f64", **{k: f[k] for k in ("lp", "li", "lv", "up", "ui", "uv")})
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            flag, x, iters, res, _ = O.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, tol, pc)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
        what = f"oracle port of GMRES(32)+ILU(0) (gmres.cu:1904-2066) double, {iters} iterations"
    ms = 1e3 * float(np.mean(times)) * scale
    sample = f"{what}; full DC solve on the {nx}x{nx} grid (n={n})" + (
        f", time scaled by n_full/n = {scale:.2f} (iteration count is size-independent for this recipe)" if scale != 1.0 else "")
    return {"value": ms, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "setup_s": setup_s, "iters": int(iters),
            "host_cores_online": os.cpu_count()}


def run_reference(args, rank):
    if rank != 0:
        return
    cb = cpu_reference(args.nx, args.cpu_nx, args.steps, args.warmup, args.tol)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["value"], "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{workload_name(args.nx)} DC solve, GMRES(32), rel. residual {args.tol:g}", "n": args.nx * args.nx},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


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
        # rank 0 must print ONE JSON line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
        os.environ["NCCL_DEBUG"] = os.environ.get("BG_NCCL_DEBUG", "WARN")

    import torch
    import torch.distributed as dist

    import b200gmres as bg

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    if world > 1:
        run_dist(args, rank, world, local)
        return
    line = run_single(args, local, args.nx, args.precond, args.steps, headline=True)
    if not args.no_configs1 and args.nx != 1304:
        # BASELINE configs[1] next to the headline: ibmpg6-sized grid, GMRES(32)+ILUT on 1 B200
        sec = run_single(args, local, 1304, "ilut", min(args.steps, 10), headline=False)
        line["configs1"] = {k: sec[k] for k in ("value", "unit", "iters", "ms_per_iteration", "true_rel_residual", "iteration_roofline", "kernels", "e2e", "setup_s")}
        line["configs1"]["config"] = sec["config"]
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(args.nx, args.cpu_nx, 1, 0, args.tol)
    print(json.dumps(line))


def run_single(args, local, nx, precond, steps, headline):
    import torch

    import b200gmres as bg

    rank, world = 0, 1

    def barrier():
        pass

    rp, ci, v, b = workload(nx)
    n = len(rp) - 1
    ctx = bg.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
    fdt = bg.F32 if args.factor_dtype == "f32" else bg.F64
    t_setup = time.perf_counter()
    if args.schedule == "block":
        o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=args.tile * args.tile_y, factor_dtype=fdt,
                            block_method=bg.PART_GRID2D if args.blocks == "grid" else bg.PART_METIS,
                            dims=[nx, nx, 1, args.tile, args.tile_y, 1], ilut_tau=args.ilut_tau, ilut_lfil=args.ilut_lfil)
    else:
        o = bg.precond_opts(trsv_schedule=bg.TRSV_LEVEL, factor_dtype=fdt, ilut_tau=args.ilut_tau, ilut_lfil=args.ilut_lfil)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o, ilut=(precond == "ilut"))
    t_setup = time.perf_counter() - t_setup
    S = bg.Solver(ctx, n, 32, dtype=bg.F64)
    orth = {"mgs": bg.ORTH_MGS, "cgs": bg.ORTH_CGS_ADAPTIVE, "cgs2": bg.ORTH_CGS2}[args.orth]
    opts = bg.gmres_opts(m=32, tol=args.tol, max_iter=10000, orth=orth, reorth_eta=args.eta)

    db = torch.from_numpy(b).cuda()
    dx = torch.zeros(n, dtype=torch.float64, device="cuda")
    hb = torch.from_numpy(b).pin_memory()
    hx = torch.zeros(n, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()

    def step_device():
        with torch.cuda.stream(stream):
            dx.zero_()  # DC solve starts from x = 0 (mna_solve_gpu_gmres.cpp:525-527)
        return S.solve(A, M, db.data_ptr(), dx.data_ptr(), opts)

    def step_host():
        hx.zero_()
        return S.solve(A, M, hb.numpy(), hx.numpy(), opts, host=True)

    def timed(fn, k):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record(stream)
        for _ in range(k):
            r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r, ctx.launches - l0

    sampler = ClockSampler(local)
    if headline:
        sampler.start()
        time.sleep(1.0)  # nvidia-smi needs a moment before its first sample
    for _ in range(max(args.warmup, 3)):
        step_device()
    tc0 = sampler.mark()
    ms, r, launches = timed(step_device, steps)
    tc1 = sampler.mark()
    clocks = sampler.stop(tc0, tc1) if headline else None
    for _ in range(2):
        step_host()
    ms_e2e, r_e2e, _ = timed(step_host, steps)

    # profiled solve: CUDA-event pairs around every launch group, on the launching stream
    with torch.cuda.stream(stream):
        dx.zero_()
    rprof, st = S.profile(A, M, db.data_ptr(), dx.data_ptr(), opts)
    x = dx.cpu().numpy()
    true_rel = float(np.linalg.norm(b - _spmv_host(rp, ci, v, x)) / np.linalg.norm(b))

    if rank == 0:
        peak, which = peaks()
        info, pinfo = A.info(), M.info()
        trsv_kernel = "trsv_block_fused (L and U solve of every diagonal block, one launch)" if args.schedule == "block" else "trsv_level_walk/_wide (L and U solve)"
        kernels = {
            "spmv_sell_w<double,5> (sliced ELL, fixed width, all loads in flight)" if info.format == bg.SPMV_SELL else "spmv_csr_vector": (st.spmv_ms, st.spmv_bytes, st.spmv_calls),
            trsv_kernel: (st.trsv_ms, st.trsv_bytes, st.trsv_calls),
            "dots_kernel (V^T w and w.w, TMA-staged tiles)": (st.dots_ms, st.dots_bytes, st.dots_calls),
            "update_kernel (w - V h, normalised, TMA-staged tiles)": (st.update_ms, st.update_bytes, st.update_calls),
        }
        if args.orth != "mgs" and st.dots_calls == 0:
            # fused step: one cooperative launch per inner iteration (projections + Givens epilogue + v_j)
            kernels = {k: v for k, v in kernels.items() if v[2] > 0}
            kernels["orth_kernel (fused: V^T w, w.w, Givens, v_j = (w - V h)/||.||; TMA-staged tiles reused from shared memory / L2)"] = (
                st.orth_ms, st.dots_bytes + st.update_bytes, st.orth_calls)
        if args.orth == "mgs":
            kernels = {k: v for k, v in kernels.items() if v[2] > 0}
            kernels["dots_kernel+update_kernel (strict MGS, i+1 pairs)"] = (st.orth_ms, st.orth_bytes, st.orth_calls)
        tot_ms = st.spmv_ms + st.trsv_ms + st.orth_ms + st.other_ms
        dom = max(kernels, key=lambda k: kernels[k][0])
        dms, dbytes, dcalls = kernels[dom]
        ach = dbytes / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
        traffic = ncu_traffic(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "frac_of_nominal_7700": ach / 7700.0,
                    "traffic": traffic.get("bytes") if traffic else None, "traffic_launch": traffic,
                    "peak_source": which, "bytes_per_launch": dbytes / max(dcalls, 1), "us_per_launch": 1e3 * dms / max(dcalls, 1),
                    "launches_timed": int(dcalls), "share_of_step": dms / tot_ms if tot_ms else None,
                    "how": "CUDA-event pairs on the library stream around every launch of the profiled solve (bg_solver_profile)"}
        all_bytes = st.spmv_bytes + st.trsv_bytes + st.orth_bytes + st.other_bytes
        per_kernel = {k: {"ms": p[0], "launches": int(p[2]), "GBps": (p[1] / (p[0] * 1e-3) / 1e9 if p[0] > 0 else 0.0),
                          "frac": (p[1] / (p[0] * 1e-3) / 1e9 / peak if p[0] > 0 else 0.0), "share": p[0] / tot_ms if tot_ms else None}
                      for k, p in kernels.items()}
        step_ms = ms / steps
        line = {
            "metric": METRIC, "value": step_ms, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{workload_name(nx)} DC solve, GMRES(32), rel. residual {args.tol:g}",
                       "n": n, "nnz": int(info.nnz), "precond": precond, "trsv_schedule": args.schedule,
                       "blocks": pinfo.nblocks, "block_rows": pinfo.max_block_rows, "factor_dtype": args.factor_dtype,
                       "orth": args.orth, "reorth_eta": args.eta, "spmv_format": "sell" if info.format == bg.SPMV_SELL else "csr",
                       "l2_policy": "working set (Krylov basis 33 x n x 8 B = %.0f MB) exceeds the 126 MB L2" % (33 * n * 8 / 1e6)},
            "iters": int(r.iters), "resid": float(r.resid), "true_rel_residual": true_rel, "reorths": int(r.reorths),
            "ms_per_iteration": step_ms / max(r.iters_done, 1),
            "iteration_roofline": {"bytes_per_iteration": all_bytes / max(rprof.iters_done, 1), "achieved_GBps": all_bytes / (step_ms * 1e-3) / 1e9,
                                   "frac": all_bytes / (step_ms * 1e-3) / 1e9 / peak,
                                   "frac_of_nominal_7700": all_bytes / (step_ms * 1e-3) / 1e9 / 7700.0,
                                   "bytes": "SURVEY.md 8(d): B_spmv + B_trsv + B_orth(j) per inner iteration, restart terms per cycle",
                                   "note": "peak = measured COPY bandwidth (read+write); read-dominated kernels can exceed it, and the "
                                           "triangular solve moves fewer bytes than the algorithmic count (16-bit local column ids), so frac may pass 1"},
            "kernels": per_kernel, "setup_s": t_setup,
            "roofline": roofline,
            "e2e": {"value": ms_e2e / steps, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": n * 8},
            "gpu_launches": int(launches), "clocks": clocks,
        }
    S.close(); M.close(); A.close(); ctx.close()
    del db, dx, hb, hx
    torch.cuda.empty_cache()
    return line


def run_dist(args, rank, world, local):
    """N > 1: the same DC solve, row-partitioned into `world` horizontal strips of the grid (contiguous
    row blocks, src/partition.cpp's initial assignment), one process per GPU: NCCL halo exchange per
    SpMV, one NCCL allreduce per orthogonalisation step, per-rank block-Jacobi ILU (dist.cu)."""
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
        # SURVEY 8(d) bytes per inner iteration summed over all ranks (the halo adds nothing to HBM traffic worth counting)
        it = max(r.iters_done, 1)
        orth_b = sum((2 * (t % 32 + 1) + 5) * s8 * n for t in range(it))
        spmv_b = ((s8 + 4) * nnz + 4 * (n + 1) + 2 * s8 * n) * it
        trsv_b = ((s8 + 4) * (info.precond.nnzL + info.precond.nnzU) * world + 8 * (n + 1) + 4 * s8 * n) * it  # nnz of rank 0 x world
        all_b = orth_b + spmv_b + trsv_b
        line = {
            "metric": METRIC, "value": step_ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{workload_name(nx)} DC solve, GMRES(32), rel. residual {args.tol:g}",
                       "n": n, "nnz": nnz, "precond": args.precond, "trsv_schedule": "block",
                       "partition": f"{world} horizontal strips (contiguous rows), per-rank block-Jacobi ILU, tiles {args.tile}x{args.tile_y}",
                       "orth": args.orth, "reorth_eta": args.eta,
                       "l2_policy": "per-rank Krylov basis 33 x n/N x 8 B = %.0f MB" % (33 * n * 8 / 1e6 / world)},
            "iters": int(r.iters), "resid": float(r.resid), "true_rel_residual": true_rel, "reorths": int(r.reorths),
            "ms_per_iteration": step_ms / it,
            "iteration_roofline": {"bytes_per_iteration": all_b / it, "achieved_GBps": all_b / (step_ms * 1e-3) / 1e9,
                                   "frac": all_b / (step_ms * 1e-3) / 1e9 / (peak * world), "peak": peak * world,
                                   "bytes": "SURVEY.md 8(d) per inner iteration, all ranks; peak = N x measured HBM copy bandwidth"},
            "comm": {"halo_bytes_per_exchange_rank0": int(info.halo_bytes_per_exchange), "neighbors_rank0": int(info.n_neighbors),
                     "allreduce_doubles_max": 34,
                     "backend": ("peer-memory kernels over NVLink (CUDA IPC windows): all-reduce fused into dots_kernel, halo = one gather+peer-store kernel"
                                 if info.transport == bg.TRANSPORT_P2P else "NCCL send/recv + allreduce on the solver stream")},
            "phases_rank0": phases, "comm_us_per_op_rank0": comm_us,
            "roofline": None, "setup_s": t_setup,
            "e2e": {"value": ms_e2e / args.steps, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": n * 8},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    D.close(); plan.close(); ctx.close()
    dist.destroy_process_group()


def ncu_traffic(kernel):
    """DRAM bytes of one captured launch of the dominant kernel (ncu --set full, read offline into
    profiles/ncu_traffic.json by tools/ncu_summary.py --traffic); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    for key, val in json.load(open(p)).items():
        if kernel.startswith(key) or key in kernel:
            return val
    return None


def _spmv_host(rp, ci, v, x):
    import scipy.sparse as sp

    return sp.csr_matrix((v, ci, rp), shape=(len(rp) - 1, len(rp) - 1)) @ x


if __name__ == "__main__":
    main()
