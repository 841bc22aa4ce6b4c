"""Host-side multi-rank logic on CPU, world_size 2 and 3: the block partition of the instances over ranks, the
padded plan an equal-sized collective needs, and the control plane that hands round the NCCL id and the CUDA
IPC handles - the library's own TCP rendezvous and its adapter over a torch.distributed (gloo) group.
The data plane (NCCL / peer-mapped slabs) needs GPUs: tests/test_gpu_multirank.py."""
import multiprocessing as mp
import os
import sys

import numpy as np

from trep_b200 import dist as D
from trep_b200 import discopt

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_range_partitions_exactly():
    assert discopt.shard_range is D.shard_range
    for n in (0, 1, 7, 8, 1000, 4096 * 9999):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
            ranges, width = D.gather_plan(n, world)
            assert ranges == r and width == max(sizes)


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rdv_worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(HERE))
    from trep_b200 import dist
    r = dist.Rendezvous(rank, world, "127.0.0.1", port, timeout=60)
    a = r.allgather(("id-from-%d" % rank).encode() * (rank + 1))          # ragged blob sizes
    r.barrier()
    b = r.allgather(b"" if rank else bytes(range(128)))                    # the unique-id pattern: only rank 0 has it
    # the partition every rank computes must tile the batch
    lo, hi = dist.shard_range(37, rank, world)
    c = r.allgather(np.array([lo, hi], np.int64).tobytes())
    r.close()
    q.put((rank, a, b, c))


def test_tcp_rendezvous_three_ranks():
    world, port = 3, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rdv_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, b, c in got:
        assert a == [("id-from-%d" % r).encode() * (r + 1) for r in range(world)]
        assert b[0] == bytes(range(128)) and b[1] == b""
        rr = [tuple(np.frombuffer(x, np.int64)) for x in c]
        assert rr[0][0] == 0 and rr[-1][1] == 37 and all(x[1] == y[0] for x, y in zip(rr, rr[1:]))


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from trep_b200 import dist as D_
    ex = D_.TorchExchange(dist)
    assert (ex.rank, ex.world) == (rank, world)
    a = ex.allgather(b"x" * (rank + 3))
    ex.barrier()
    g = D_.Group(device=0, exchange=ex)         # no communicator is created until a data-plane call needs it
    assert g._comm is None and (g.rank, g.world) == (rank, world)
    q.put((rank, a))
    dist.barrier()
    dist.destroy_process_group()


def test_torch_exchange_over_gloo_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] == [b"xxx", b"xxxx"]


def test_single_rank_group_needs_no_network():
    g = D.Group(device=0, exchange=D.Rendezvous(0, 1))
    assert g.exchange.allgather(b"abc") == [b"abc"]
    g.barrier()
    g.close()
