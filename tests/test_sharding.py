"""Host-side multi-rank logic on CPU (gloo, world_size 2): block partition of the instances over
ranks and the all-gather of the A/B slabs (the only exchange of the path).
The per-rank compute is injected (tests/hostmath.py, test infrastructure) because this container
has no GPU; on the GPU box the same code path runs the CUDA library per rank."""
import os
import sys
import types

import numpy as np
import pytest

import golden_util as G
from trep_b200 import discopt

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1000, 4096 * 9999):
        for world in (1, 2, 3, 8):
            r = [discopt.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _host_compute(desc):
    import hostmath as H

    def fn(q1, p1, u1, rho2, t1, t2, hint):
        n = q1.shape[0]
        A = np.zeros((n, desc.nX, desc.nX)); B = np.zeros((n, desc.nX, desc.nU)); st = np.zeros(n, np.int32)
        for i in range(n):
            o = H.linearize(desc, t1[i], t2[i], q1[i], p1[i], u1[i], rho2[i], q2_guess=None if hint is None else hint[i])
            A[i], B[i], st[i] = o["A"], o["B"], o["rc"]
        return A, B, st
    return fn


def _inputs(desc, n):
    rng = np.random.default_rng(5)
    X = np.concatenate([rng.uniform(-1, 1, (n, desc.nq)), rng.normal(0, 1, (n, desc.nd)), np.zeros((n, desc.nk))], axis=1)
    U = rng.uniform(-1, 1, (n, desc.nU))
    t1 = 0.01 * np.arange(n)
    return X, U, t1, t1 + 0.01


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = G.desc("pend_on_cart1")
    v = types.SimpleNamespace(nq=desc.nq, nd=desc.nd, nk=desc.nk, nu=desc.nu, tolerance=1e-10)
    ds = discopt.DSystem(v, np.arange(0, 1, 0.01))
    X, U, t1, t2 = _inputs(desc, 37)       # ragged: 19 + 18
    A, B = ds.linearize(X, U, t1, t2, dist=dist, compute=_host_compute(desc))
    if rank == 0:
        q.put((A, B))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    A, B = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    desc = G.desc("pend_on_cart1")
    v = types.SimpleNamespace(nq=desc.nq, nd=desc.nd, nk=desc.nk, nu=desc.nu, tolerance=1e-10)
    ds = discopt.DSystem(v, np.arange(0, 1, 0.01))
    X, U, t1, t2 = _inputs(desc, 37)
    A1, B1 = ds.linearize(X, U, t1, t2, compute=_host_compute(desc))
    assert np.array_equal(A, A1) and np.array_equal(B, B1)


def test_linearize_trajectory_shapes_and_layout():
    desc = G.desc("pend_on_cart1")
    v = types.SimpleNamespace(nq=desc.nq, nd=desc.nd, nk=desc.nk, nu=desc.nu, tolerance=1e-10)
    t = np.arange(0, 0.1, 0.01)
    ds = discopt.DSystem(v, t)
    rng = np.random.default_rng(0)
    X = rng.normal(0, 0.3, (2, 10, ds.nX)); U = rng.normal(0, 1, (2, 9, ds.nU))
    A, B = ds.linearize_trajectory(X, U, compute=_host_compute(desc))
    assert A.shape == (2, 9, 4, 4) and B.shape == (2, 9, 4, 1)
    A0, B0 = ds.linearize_trajectory(X[1], U[1], compute=_host_compute(desc))
    assert np.array_equal(A0, A[1]) and np.array_equal(B0, B[1])
    # state packing round trip
    Q, p, vv = ds.split_state(X)
    assert np.array_equal(ds.build_state(Q, p, vv), X)


# ---- Monte-Carlo sweep (BASELINE config 4): rollouts block-partitioned over ranks, final states gathered
def _host_sweep(desc, dt, nsteps):
    import hostmath as H

    def fn(q0, q1, us, ks):
        n = q0.shape[0]
        q2 = np.zeros((n, desc.nq)); p2 = np.zeros((n, desc.nd)); it = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
        for i in range(n):
            p0 = H.calc_p2(desc, dt, q0[i], q1[i])
            rc, q2[i], p2[i], lam, it[i] = H.step(desc, nsteps, dt, dt, q1[i], p0)
            st[i] = rc
        return q2, p2, it, st
    return fn


def _sweep_inputs():
    return np.random.default_rng(8).uniform(-np.pi, np.pi, (11, 2))      # ragged over 2 ranks: 6 + 5


def _sweep_worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from trep_b200.midpointvi import monte_carlo_sweep
    desc = G.desc("dual_pendulums")
    out = monte_carlo_sweep(desc, _sweep_inputs(), 0.01, 40, dist=dist, compute=_host_sweep(desc, 0.01, 40))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_monte_carlo_sweep_two_ranks_matches_single_process():
    import torch.multiprocessing as mp
    from trep_b200.midpointvi import monte_carlo_sweep
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    desc = G.desc("dual_pendulums")
    want = monte_carlo_sweep(desc, _sweep_inputs(), 0.01, 40, compute=_host_sweep(desc, 0.01, 40))
    for k in ("q2", "p2", "iters", "status", "hist"):
        assert np.array_equal(got[k], want[k]), k
    assert want["status"].tolist() == [0] * 11 and want["hist"].sum() == 11
