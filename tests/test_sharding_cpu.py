"""World-size-2 gloo test of the multi-GPU host logic: buoys are partitioned over ranks, every rank deposits
into a private accumulator [b | misfit | n_masked], one all-reduce(sum) makes them identical everywhere and equal
to the single-rank result.  The per-rank arithmetic here is the CPU oracle - this tests the
partition / reduction plumbing (ocp_b200.sharding), not the kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from ocp_b200.sharding import allreduce_accumulator, shard_bounds, shard_buoys


def test_shard_bounds_cover_and_balance():
    for K in (0, 1, 7, 10, 10000, 10**7 + 3):
        for n in (1, 2, 3, 8):
            b = [shard_bounds(K, r, n) for r in range(n)]
            assert b[0][0] == 0 and b[-1][1] == K
            assert all(b[i][1] == b[i + 1][0] for i in range(n - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, H.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V = H.square32()
        xr, ud = H.traj(10)
        x0, ud_loc = shard_buoys(xr[:, 0, :], ud, rank, world)
        P = H.OraclePipeline(V, 1.0, x0, ud_loc, 1e-5)
        w = H.field_for(100)
        g = P.O.project_gradient(w)
        x, u, cell, mask, parked = P.primal(w)
        mu = P.B.adjoint(g, x, u, ud_loc, mask, H.H)
        bn = P.B.point_sources(V.velocity_nodal(w), x, ud_loc, mu, mask, H.H, H.CENTER)
        acc = torch.from_numpy(np.r_[bn.ravel(), P.B.misfit(u, ud_loc, H.H), mask.sum()])
        allreduce_accumulator(acc, dist.group.WORLD)
        q.put((rank, acc.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_rank():
    V = H.square32()
    xr, ud = H.traj(10)
    P = H.OraclePipeline(V, 1.0, xr[:, 0, :].copy(), ud, 1e-5)
    w = H.field_for(100)
    g = P.O.project_gradient(w)
    x, u, cell, mask, parked = P.primal(w)
    mu = P.B.adjoint(g, x, u, ud, mask, H.H)
    bn = P.B.point_sources(V.velocity_nodal(w), x, ud, mu, mask, H.H, H.CENTER)
    ref = np.r_[bn.ravel(), P.B.misfit(u, ud, H.H), mask.sum()]

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0], got[1])                       # identical on every rank after the all-reduce
    assert H.rel(got[0], ref) < 1e-12                           # summation order differs from the 1-rank sum


def test_spatial_order_is_a_permutation_grouping_neighbours():
    from ocp_b200.sharding import spatial_order
    V = H.square32()
    rng = np.random.default_rng(0)
    x0 = np.stack([rng.uniform(-0.1, 2.1, 5000), rng.uniform(-0.1, 2.1, 5000)], 1)
    x0[17] = np.nan
    p = spatial_order(V, x0)
    assert sorted(p.tolist()) == list(range(5000))
    xs = x0[p]
    ok = ~np.isnan(xs).any(axis=1)
    d_sorted = np.linalg.norm(np.diff(xs[ok], axis=0), axis=1).mean()
    d_orig = np.linalg.norm(np.diff(x0[~np.isnan(x0).any(axis=1)], axis=0), axis=1).mean()
    assert d_sorted < 0.1 * d_orig                     # consecutive buoys are neighbours after sorting
