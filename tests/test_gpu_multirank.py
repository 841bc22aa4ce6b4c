"""Multi-rank data plane on real GPUs (trep_b200/dist.py over the C ABI's trepb_comm_* / trepb_ipc_*): one
process per rank.  The peer-mapped (CUDA IPC) gather runs with both ranks on ONE GPU, so it is covered on the
single-GPU test box; the NCCL gather needs one GPU per rank and is skipped below two devices (the bench runs it
at N = 2, 4, 8)."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(name, R, K):
    rng = np.random.default_rng(11)
    d = G.desc(name)
    t = 0.01 * np.arange(K + 1)
    X = np.zeros((R, K + 1, d.nX)); U = np.zeros((R, K, d.nU))
    X[..., :d.nq] = rng.uniform(-0.5, 0.5, (R, K + 1, d.nq)); X[..., d.nq:d.nq + d.nd] = rng.normal(0, 1, (R, K + 1, d.nd))
    U[:] = rng.uniform(-1, 1, U.shape)
    return t, X, U


def _worker(rank, world, port, name, gather, one_gpu, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    from trep_b200 import discopt, dist, midpointvi
    dev = 0 if one_gpu else rank
    ex = dist.Rendezvous(rank, world, "127.0.0.1", port, timeout=120)
    grp = dist.Group(device=dev, exchange=ex)
    t, X, U = _inputs(name, 3, 13)          # 39 linearizations: ragged over 2 ranks (20 + 19)
    ds = discopt.DSystem(midpointvi.MidpointVI(G.desc(name), device=dev), t)
    res = ds.linearize_trajectory(X, U, group=grp, gather=gather)
    if rank == 0:
        q.put((res.A, res.B))
    else:
        assert res is None
    grp.barrier()
    grp.close()


def _run(name, gather, one_gpu):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, gather, one_gpu, q)) for r in range(2)]
    for p in procs:
        p.start()
    A, B = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from trep_b200 import discopt, midpointvi
    t, X, U = _inputs(name, 3, 13)
    ds = discopt.DSystem(midpointvi.MidpointVI(G.desc(name)), t)
    want = ds.linearize_trajectory(X, U)
    assert np.array_equal(A, want.A) and np.array_equal(B, want.B)


def test_peer_mapped_gather_two_ranks_on_one_gpu():
    """gather="peer": both ranks' linearize kernels write their blocks straight into rank 0's slab (CUDA IPC)."""
    _run("pend_on_cart1", "peer", True)


def test_nccl_gather_two_ranks():
    from trep_b200 import lib
    if lib.device_count() < 2:
        pytest.skip("NCCL needs one GPU per rank")
    _run("pend_on_cart1", "nccl", False)
    _run("pend_on_cart1", "peer", False)


def _sweep_worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    from trep_b200 import dist
    from trep_b200.midpointvi import monte_carlo_sweep
    grp = dist.Group(device=rank, exchange=dist.Rendezvous(rank, world, "127.0.0.1", port, timeout=120))
    q0 = np.random.default_rng(8).uniform(-np.pi, np.pi, (1001, 2))
    out = monte_carlo_sweep(G.desc("dual_pendulums"), q0, 0.01, 40, group=grp)
    if rank == 0:
        q.put(out)
    grp.barrier()
    grp.close()


def test_monte_carlo_sweep_two_ranks_matches_single_process():
    from trep_b200 import lib
    from trep_b200.midpointvi import monte_carlo_sweep
    if lib.device_count() < 2:
        pytest.skip("NCCL needs one GPU per rank")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    q0 = np.random.default_rng(8).uniform(-np.pi, np.pi, (1001, 2))
    want = monte_carlo_sweep(G.desc("dual_pendulums"), q0, 0.01, 40)
    for k in ("q2", "p2", "iters", "status", "hist"):
        assert np.array_equal(got[k], want[k]), k
