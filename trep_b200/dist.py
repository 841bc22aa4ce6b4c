"""Multi-GPU plumbing of the batched path: one process per GPU, no torch.

The batch shards over ranks with no exchange during compute; what this module provides is

  shard_range / gather_plan   the contiguous block partition of n instances over the ranks
  Rendezvous                  a small TCP control plane (rank 0 listens on MASTER_ADDR:port): all-gather of
                              byte blobs and a barrier - enough to hand round the 128-byte NCCL id and the
                              64-byte CUDA IPC handles; `TorchExchange` offers the same two calls on top of
                              an initialised torch.distributed group for hosts that already have one
  Comm                        NCCL communicator behind the C ABI (trepb_comm_*): all-gather / gather of
                              device-resident slabs over NVLink
  SharedSlab                  one rank's HBM slab mapped into every rank (trepb_ipc_*): a rank that passes
                              `slab.ptr + its offset` as the A / B outputs of trepb_linearize_batch_dev makes the
                              linearize kernel itself deliver its results into the root's memory
  gather_rows                 [n_local, ...] slabs in HBM -> [n_total, ...] on the root (or on every rank)

Replaces the round-1 route (device -> host -> torch tensor -> device -> NCCL -> host).
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import struct
import time

import numpy as np


def shard_range(n, rank, world):
    """Contiguous block partition of n instances: [lo, hi) of `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_plan(n, world):
    """(ranges, width): every rank's [lo, hi) and the padded per-rank row count an equal-sized
    collective (ncclAllGather) needs."""
    ranges = [shard_range(n, r, world) for r in range(world)]
    return ranges, max(hi - lo for lo, hi in ranges) if ranges else 0


# ---- control plane ------------------------------------------------------------------------------------
def _send_blob(sock, blob):
    sock.sendall(struct.pack("!Q", len(blob)) + blob)


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed the rendezvous connection")
        buf += chunk
    return bytes(buf)


def _recv_blob(sock):
    (n,) = struct.unpack("!Q", _recv_exact(sock, 8))
    return _recv_exact(sock, n)


class Rendezvous:
    """All-gather of byte blobs + barrier over plain TCP; rank 0 is the hub.  Addresses default to the
    torchrun environment (MASTER_ADDR, MASTER_PORT + 1 so that a torch store on MASTER_PORT is left alone)."""

    def __init__(self, rank=None, world=None, addr=None, port=None, timeout=120.0):
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(port if port is not None else int(os.environ.get("MASTER_PORT", "29500")) + 1)
        self.peers = []
        self.sock = None
        if self.world == 1:
            return
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(self.world)
            srv.settimeout(timeout)
            got = {}
            while len(got) < self.world - 1:
                c, _ = srv.accept()
                c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                c.settimeout(timeout)
                (r,) = struct.unpack("!I", _recv_exact(c, 4))
                got[r] = c
            srv.close()
            self.peers = [got[r] for r in range(1, self.world)]
        else:
            deadline = time.time() + timeout
            while True:
                try:
                    s = socket.create_connection((addr, port), timeout=timeout)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            s.settimeout(timeout)
            s.sendall(struct.pack("!I", self.rank))
            self.sock = s

    def allgather(self, blob: bytes):
        """Every rank's blob, in rank order, on every rank."""
        blob = bytes(blob)
        if self.world == 1:
            return [blob]
        if self.rank == 0:
            parts = [blob] + [_recv_blob(c) for c in self.peers]
            packed = b"".join(struct.pack("!Q", len(p)) + p for p in parts)
            for c in self.peers:
                _send_blob(c, packed)
            return parts
        _send_blob(self.sock, blob)
        packed = _recv_blob(self.sock)
        parts, o = [], 0
        for _ in range(self.world):
            (n,) = struct.unpack("!Q", packed[o:o + 8])
            parts.append(packed[o + 8:o + 8 + n])
            o += 8 + n
        return parts

    def barrier(self):
        self.allgather(b"")

    def close(self):
        for c in self.peers:
            c.close()
        if self.sock is not None:
            self.sock.close()
        self.peers, self.sock = [], None


class TorchExchange:
    """The Rendezvous interface (rank, world, allgather, barrier) on an initialised torch.distributed group
    (gloo or nccl) - for hosts that run under torchrun anyway.  torch is imported by the caller, not here."""

    def __init__(self, dist):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allgather(self, blob: bytes):
        out = [None] * self.world
        self.dist.all_gather_object(out, bytes(blob))
        return out

    def barrier(self):
        self.dist.barrier()

    def close(self):
        pass


# ---- data plane (needs the CUDA library) ----------------------------------------------------------------
def _lib():
    from . import lib
    return lib


class Comm:
    """NCCL communicator of the C ABI (include/trepb.h: trepb_comm_*), one rank per GPU."""

    def __init__(self, device, exchange):
        L = _lib()
        raw = L.raw()
        self.device, self.rank, self.world = device, exchange.rank, exchange.world
        ident = C.create_string_buffer(128)
        if self.rank == 0:
            L._check(raw.trepb_comm_unique_id(ident))
        ident = C.create_string_buffer(exchange.allgather(ident.raw if self.rank == 0 else b"")[0], 128)
        h = C.c_void_p()
        L._check(raw.trepb_comm_create(device, self.rank, self.world, ident, C.byref(h)))
        self._h = h

    def allgather(self, send, recv, bytes_per_rank, stream=None):
        L = _lib()
        L._check(L.raw().trepb_comm_allgather_dev(self._h, C.c_void_p(L._ptr(send)), C.c_void_p(L._ptr(recv)),
                                                  C.c_int64(bytes_per_rank), stream))

    def gather(self, send, recv, bytes_per_rank, root=0, stream=None):
        L = _lib()
        L._check(L.raw().trepb_comm_gather_dev(self._h, C.c_void_p(L._ptr(send)), C.c_void_p(L._ptr(recv)),
                                               C.c_int64(bytes_per_rank), root, stream))

    def close(self):
        if getattr(self, "_h", None):
            L = _lib()
            L.raw().trepb_comm_destroy.restype = None
            L.raw().trepb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Addr:
    """pointer-like (data_ptr()) at a fixed device address"""

    def __init__(self, p):
        self.p = int(p)

    def data_ptr(self):
        return self.p


class SharedSlab:
    """`nbytes` of the root's HBM visible on every rank (CUDA IPC, peer access over NVLink).  `.at(offset)` is
    a pointer-like for the raw entry points; on the root it is the local allocation itself."""

    def __init__(self, device, exchange, nbytes, root=0):
        L = _lib()
        raw = L.raw()
        self.device, self.rank, self.world, self.root, self.nbytes = device, exchange.rank, exchange.world, root, int(nbytes)
        self.local = None
        self._mapped = None
        handle = b""
        if self.rank == root:
            self.local = L.DeviceBuffer(device, (max(self.nbytes, 8),), np.uint8)
            if self.world > 1:
                hb = C.create_string_buffer(64)
                L._check(raw.trepb_ipc_export(device, C.c_void_p(self.local.ptr), hb))
                handle = hb.raw
        handle = exchange.allgather(handle)[root]
        if self.rank == root:
            self.ptr = self.local.ptr
        else:
            p = C.c_void_p()
            L._check(raw.trepb_ipc_open(device, C.create_string_buffer(handle, 64), C.byref(p)))
            self._mapped = p.value
            self.ptr = p.value

    def at(self, offset_bytes):
        return _Addr(self.ptr + int(offset_bytes))

    def close(self):
        L = _lib()
        if self._mapped:
            L.raw().trepb_ipc_close(self.device, C.c_void_p(self._mapped))
            self._mapped = None
        if self.local is not None:
            self.local.free()
            self.local = None


class Group:
    """What a multi-rank call needs: the control plane (`exchange`: Rendezvous or TorchExchange), this rank's
    device and - created on first use - the NCCL communicator."""

    def __init__(self, device=None, exchange=None):
        self.exchange = exchange if exchange is not None else Rendezvous()
        self.rank, self.world = self.exchange.rank, self.exchange.world
        self.device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
        self._comm = None

    @property
    def comm(self):
        if self._comm is None:
            self._comm = Comm(self.device, self.exchange)
        return self._comm

    def barrier(self):
        self.exchange.barrier()

    def close(self):
        if self._comm is not None:
            self._comm.close()
        self.exchange.close()


def gather_rows(local, n_local, n_total, row_bytes, comm: Comm, device, everywhere=False, stream=None):
    """Concatenate per-rank device slabs (`local`: pointer-like, n_local rows of row_bytes) in rank order.
    Returns a DeviceBuffer of n_total rows on the root (rank 0) - on every rank with everywhere=True - else
    None.  Ragged partitions are padded to the widest shard for the collective and compacted on the device."""
    L = _lib()
    ranges, width = gather_plan(n_total, comm.world)
    ragged = any(hi - lo != width for lo, hi in ranges)
    send = local
    pad = None
    if ragged:
        pad = L.DeviceBuffer(device, (width * row_bytes,), np.uint8)
        L._check(L.raw().trepb_memcpy_d2d(device, C.c_void_p(pad.ptr), C.c_void_p(L._ptr(local)), C.c_int64(n_local * row_bytes)))
        send = pad
    want = everywhere or comm.rank == 0
    full = L.DeviceBuffer(device, (comm.world * width * row_bytes,), np.uint8) if want else None
    if everywhere:
        comm.allgather(send, full, width * row_bytes, stream)
    else:
        comm.gather(send, full, width * row_bytes, 0, stream)
    L.synchronize(device)
    if pad is not None:
        pad.free()
    if not want:
        return None
    if not ragged:
        return full
    out = L.DeviceBuffer(device, (n_total * row_bytes,), np.uint8)
    for r, (lo, hi) in enumerate(ranges):
        if hi > lo:
            L._check(L.raw().trepb_memcpy_d2d(device, C.c_void_p(out.ptr + lo * row_bytes),
                                              C.c_void_p(full.ptr + r * width * row_bytes), C.c_int64((hi - lo) * row_bytes)))
    full.free()
    return out
