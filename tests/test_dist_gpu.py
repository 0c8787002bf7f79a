"""Row-partitioned solve on real GPUs.

 * one rank, one GPU: the row-partitioned entry points with a trivial partition must reproduce the
   single-GPU solver exactly (same kernels, no communication);
 * two ranks (only when the box has >= 2 GPUs): tools/dist_check.py under torchrun -- both transports
   (peer-memory kernels over NVLink and the NCCL baseline) against a direct solve, same iteration count.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_dist_equals_solver(bg, ctx):
    from b200gmres import synth

    nx, ny, tile = 96, 80, 16
    rp, ci, v, b = synth.pg2d(nx, ny, seed=7)
    n = len(rp) - 1
    o = bg.precond_opts(trsv_schedule=bg.TRSV_BLOCK, block_rows=tile * tile, block_method=bg.PART_GRID2D, dims=[nx, ny, 1, tile, tile, 1])
    opts = bg.gmres_opts(m=32, tol=1e-8, max_iter=2000, reorth_eta=0.1)
    plan = bg.DistPlan(rp, ci, v, 1, 0, np.zeros(n, np.int32))
    assert plan.info.n_local == n and plan.info.n_halo == 0 and plan.info.n_neighbors == 0
    D = bg.Dist(ctx, plan, None, dtype=bg.F64, m=32, opts=o)
    db = torch.from_numpy(b).cuda()
    dx = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    r = D.solve(db.data_ptr(), dx.data_ptr(), opts)
    ctx.sync()
    x_dist = dx.cpu().numpy()
    A = bg.Matrix(ctx, rp, ci, v, dtype=bg.F64)
    M = bg.Precond.ilu0(ctx, rp, ci, v, dtype=bg.F64, opts=o)
    S = bg.Solver(ctx, n, 32, dtype=bg.F64)
    dx.zero_()
    torch.cuda.synchronize()
    r1 = S.solve(A, M, db.data_ptr(), dx.data_ptr(), opts)
    ctx.sync()
    x_one = dx.cpu().numpy()
    assert r.flag == 0 and r1.flag == 0 and r.iters == r1.iters
    assert np.array_equal(x_dist, x_one)  # same kernels in the same order: bit-identical
    # host-buffer entry point (H2D, solve, D2H)
    hx = np.zeros(n)
    rh = D.solve(b, hx, opts, host=True)
    assert rh.flag == 0 and np.array_equal(hx, x_one)
    S.close(); M.close(); A.close(); D.close(); plan.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_two_ranks_both_transports():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert '"ok": true' in p.stdout and '"ok": false' not in p.stdout
