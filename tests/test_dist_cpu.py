"""Multi-rank host logic on CPU (world_size 2 and 3, gloo): every rank builds its own plan from the
global pattern; the halo exchange the plans describe, run with torch.distributed send/recv, gives
every rank exactly the neighbour values its local rows need, so the row-partitioned SpMV equals the
global one and the partial dots allreduce to the global dot.  (The GPU path replaces the numpy
arithmetic with kernels and gloo with NCCL; the plan arrays are the same.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, method, q):
    try:
        import torch
        import torch.distributed as dist
        sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
        import b200gmres as bg
        from b200gmres import synth
        import scipy.sparse as sp

        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        rp, ci, v, b = synth.pg2d(37, 29, seed=5) if method != "rand" else synth.random_sparse(900, 6, seed=9)
        n = len(rp) - 1
        meth = {"contig": bg.PART_CONTIGUOUS, "metis": bg.PART_METIS, "rand": bg.PART_CONTIGUOUS}[method]
        nb, part = bg.host_block_partition(meth, rp, ci, world)
        plan = bg.DistPlan(rp, ci, v, world, rank, part)
        a = plan.arrays()
        i = plan.info
        assert i.n_local == int((part == rank).sum()) and np.array_equal(a["l2g"], np.nonzero(part == rank)[0])
        assert np.all(part[a["h2g"]] != rank)
        x = np.cos(0.1 * np.arange(n)) + 2.0  # global test vector
        xl = x[a["l2g"]].copy()
        # halo exchange exactly as the plan prescribes: pack own rows, send, receive into the halo slots
        halo = np.zeros(i.n_halo)
        reqs, bufs = [], []
        for k, qr in enumerate(a["nbr"]):
            s = torch.from_numpy(xl[a["send_idx"][a["send_ptr"][k]:a["send_ptr"][k + 1]]].copy())
            r = torch.zeros(int(a["recv_ptr"][k + 1] - a["recv_ptr"][k]), dtype=torch.float64)
            bufs.append((k, r))
            if s.numel():
                reqs.append(dist.isend(s, int(qr)))
            if r.numel():
                reqs.append(dist.irecv(r, int(qr)))
        for rq in reqs:
            rq.wait()
        for k, r in bufs:
            halo[a["recv_ptr"][k]:a["recv_ptr"][k + 1]] = r.numpy()
        assert np.array_equal(halo, x[a["h2g"]])  # bit-exact: pure data movement
        Aloc = sp.csr_matrix((a["val"], a["col"], a["rowptr"]), shape=(i.n_local, i.n_local + i.n_halo))
        yl = Aloc @ np.concatenate([xl, halo])
        y = sp.csr_matrix((v, ci, rp), shape=(n, n)) @ x
        assert np.allclose(yl, y[a["l2g"]], rtol=1e-14, atol=1e-14)
        # partial dots -> allreduce
        t = torch.tensor([float(xl @ yl), float(i.n_local)], dtype=torch.float64)
        dist.all_reduce(t)
        assert abs(t[0].item() - float(x @ y)) <= 1e-12 * abs(float(x @ y)) and int(t[1].item()) == n
        plan.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("world,method", [(2, "contig"), (2, "metis"), (3, "contig"), (2, "rand")])
def test_dist_plan_halo_and_allreduce_gloo(world, method):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, method, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_dist_plan_single_rank_is_identity():
    sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
    import b200gmres as bg
    from b200gmres import synth

    rp, ci, v, b = synth.pg2d(11, 9, seed=3)
    plan = bg.DistPlan(rp, ci, v, 1, 0, np.zeros(len(rp) - 1, np.int32))
    a = plan.arrays()
    assert plan.info.n_halo == 0 and plan.info.n_neighbors == 0
    assert np.array_equal(a["rowptr"], rp) and np.array_equal(a["col"], ci) and np.array_equal(a["val"], v)
    plan.close()
