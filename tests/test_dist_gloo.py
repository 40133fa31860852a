"""CPU suite: the multi-process merge logic (density normaliser, integer grids, cost) with world_size 2 over gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from density_planner_b200 import dist as D
    import dp_oracle as O
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    rs = np.random.RandomState(0)
    N, Ts = 1001, 7
    rholog = rs.normal(0, 30, (N, Ts)).astype(np.float32)
    rho0 = rs.uniform(0.1, 1, N).astype(np.float32)
    x = rs.normal(0, 2, N).astype(np.float32)
    y = rs.normal(-10, 4, N).astype(np.float32)
    lo, hi = D.shard_range(N, rank, world)
    # normaliser: local stats -> merged stats -> local apply, compared with the single-process oracle
    l = torch.from_numpy(rholog[lo:hi])
    lmax = l.max(dim=0).values
    lsum = (torch.exp((l - lmax).double()) * torch.from_numpy(rho0[lo:hi]).double()[:, None]).sum(0)
    gmax, gsum = D.merge_normalizer_stats(lmax, lsum)
    rho_local = (torch.exp(l - gmax) * torch.from_numpy(rho0[lo:hi])[:, None] / gsum.float()).numpy()
    ref = O.normalize_density(rholog, rho0)[lo:hi]
    ok_norm = np.allclose(rho_local, ref, rtol=1e-5, atol=1e-12)
    # integer grids: per-shard binning, allreduce, equals binning of the whole set
    scale = 2.0 ** 40
    gx, gy = O.pos2gridpos(x, y)
    gx, gy = np.clip(gx, 0, 241), np.clip(gy, 0, 401)
    def bins(sl):
        m = np.zeros((242, 402), np.int64)
        c = np.zeros((242, 402), np.int32)
        np.add.at(m, (gx[sl], gy[sl]), np.rint(rho0[sl].astype(np.float64) * scale).astype(np.int64))
        np.add.at(c, (gx[sl], gy[sl]), 1)
        return torch.from_numpy(m), torch.from_numpy(c)
    m, c = bins(slice(lo, hi))
    D.allreduce_grids(m, c)
    mf, cf = bins(slice(0, N))
    ok_grid = bool((m == mf).all()) and bool((c == cf).all())
    part = torch.tensor([float(rank + 1), 2.0], dtype=torch.float64)
    D.allreduce_sum(part)
    ok_sum = part.tolist() == [3.0, 4.0]
    tmax = D.max_over_ranks(1.0 + rank, "cpu")
    q.put((rank, ok_norm, ok_grid, ok_sum, tmax, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert [r[5] for r in res] == [(0, 501), (501, 1001)]
    for r in res:
        assert r[1] and r[2] and r[3] and r[4] == 2.0, r


def test_shard_range_covers_everything():
    from density_planner_b200.dist import shard_range
    for n in (0, 1, 7, 1000, 10**8 + 3):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _pipeline_worker(rank, world, port, q):
    """The two collectives of the device-resident pipeline (density_planner_b200/pipeline.py) on emulated shard results:
    MAX of the per-slice log-density maxima, then ONE SUM over the packed int64 buffer (mass | int32 count pairs | cost)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from density_planner_b200 import dist as D
    import dp_oracle as O
    D.init_from_env(backend="gloo")
    rs = np.random.RandomState(1)
    N, Ts, cx, cy = 4001, 3, 242, 402
    l = rs.normal(0, 20, (Ts, N)).astype(np.float32)
    rho0 = rs.uniform(0.5, 1.5, N).astype(np.float32)
    x = rs.normal(0, 1.5, (Ts, N)).astype(np.float32)
    y = rs.normal(-10, 3, (Ts, N)).astype(np.float32)
    term = rs.uniform(0, 30, (Ts, N)).astype(np.float32)  # stands for p * |des - pos|^2 of the sample's cell
    ks, kc = 34, 20

    def shard_buffer(sl, lmax):
        gc = Ts * cx * cy
        buf = np.zeros(gc + (gc + 1) // 2 + Ts, np.int64)
        mass = buf[:gc].reshape(Ts, cx, cy)
        count = buf[gc:gc + (gc + 1) // 2].view(np.int32)[:gc].reshape(Ts, cx, cy)
        cost = buf[gc + (gc + 1) // 2:]
        for t in range(Ts):
            w = (np.exp((l[t, sl] - lmax[t]).astype(np.float32)).astype(np.float32) * rho0[sl]).astype(np.float32)
            gx, gy = O.pos2gridpos(x[t, sl], y[t, sl])
            gx, gy = np.clip(gx, 0, cx - 1), np.clip(gy, 0, cy - 1)
            np.add.at(mass[t], (gx, gy), np.rint(w.astype(np.float64) * 2.0 ** ks).astype(np.int64))
            np.add.at(count[t], (gx, gy), 1)
            cost[t] = np.rint((w * term[t, sl]).astype(np.float32).astype(np.float64) * 2.0 ** kc).astype(np.int64).sum()
        return buf

    lo, hi = D.shard_range(N, rank, world)
    lmax = torch.from_numpy(l[:, lo:hi].max(axis=1))
    dist.all_reduce(lmax, op=dist.ReduceOp.MAX)
    ok_max = bool((lmax.numpy() == l.max(axis=1)).all())
    buf = torch.from_numpy(shard_buffer(slice(lo, hi), lmax.numpy()))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    full = shard_buffer(slice(0, N), l.max(axis=1))
    ok_buf = bool((buf.numpy() == full).all())
    q.put((rank, ok_max, ok_buf))
    dist.barrier()
    dist.destroy_process_group()


def test_pipeline_collectives_world2_gloo():
    """Sharded pipeline result == single-process result, bit for bit: integer accumulators, a global max before the
    weights are quantised, int32 counts summed pairwise inside int64 words without carries."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] and r[2], r
