"""CPU baseline: the REFERENCE's own C implementation (oracle/_ref, unmodified numerics) timed in
a tight C loop (oracle/ref_harness.c) on the host cores.

TEST INFRASTRUCTURE ONLY - used by bench.py's `cpu_baseline` leg and `--impl reference` arm and by
tests/, never by trep_b200/.

The reference is single-threaded for the DEL step and deriv1 (GIL held, no
Py_BEGIN_ALLOW_THREADS in midpointvi.c), so all host cores are used the only way the reference
allows: one independent process per core, each with its own System/MidpointVI, on a disjoint
shard of the batch.
"""
import ctypes as C
import multiprocessing as mp
import os
import subprocess
import sys
import sysconfig
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ref_harness.c")
SO = os.path.join(HERE, "_ref", "ref_harness.so")


def build_harness(force=False):
    """gcc -O2 of oracle/ref_harness.c against the headers the reference installs."""
    hdr = os.path.join(HERE, "_ref", "trep", "_trep")
    if os.path.exists(SO) and not force and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    import numpy
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-I" + sysconfig.get_paths()["include"],
           "-I" + numpy.get_include(), "-I" + hdr, SRC, "-o", SO]
    subprocess.check_call(cmd)
    return SO


_dp = C.POINTER(C.c_double)
_ipt = C.POINTER(C.c_int)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Harness:
    def __init__(self, name, tolerance=1e-10):
        if HERE not in sys.path:
            sys.path.insert(0, HERE)
        import ref_systems as R
        self.R = R
        self.system, self.mvi = R.make_mvi(name, tolerance)
        self.h = C.PyDLL(build_harness())
        ext = C.PyDLL(R.trep._trep.__file__)
        self.solve = C.cast(ext.MidpointVI_solve_DEL, C.c_void_p)
        self.deriv1 = C.py_object("_calc_deriv1")
        m = self.mvi
        self.nq, self.nd, self.nk, self.nu, self.nc = m.nq, m.nd, m.nk, m.nu, m.nc
        self.h.rh_rollouts.restype = C.c_long
        self.h.rh_linearize.restype = C.c_long

    def rollouts(self, q, p, nsteps, t0, dt, max_it=200):
        q = np.ascontiguousarray(q, float); p = np.ascontiguousarray(p, float)
        B = q.shape[0]
        q2 = np.empty_like(q); p2 = np.empty_like(p)
        it = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
        self.h.rh_rollouts(C.c_void_p(id(self.mvi)), self.solve, C.c_long(B), C.c_int(nsteps),
                           C.c_int(self.nq), C.c_int(self.nd), C.c_int(self.nc), C.c_double(t0),
                           C.c_double(dt), C.c_int(max_it), _p(q), _p(p), _p(q2), _p(p2), _p(it), _p(st))
        return dict(q2=q2, p2=p2, iters=it, status=st)

    def linearize(self, q1, p1, u1, k2, t1, dt, q2_hint=None, lam_hint=None, max_it=200):
        q1 = np.ascontiguousarray(q1, float); p1 = np.ascontiguousarray(p1, float)
        B = q1.shape[0]
        u1 = None if not self.nu else np.ascontiguousarray(u1, float)
        k2 = None if not self.nk else np.ascontiguousarray(k2, float)
        q2_hint = None if q2_hint is None else np.ascontiguousarray(q2_hint, float)
        lam_hint = None if lam_hint is None else np.ascontiguousarray(lam_hint, float)
        nX, nU = 2 * self.nq, self.nu + self.nk
        A = np.empty((B, nX, nX)); Bm = np.empty((B, nX, nU))
        it = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
        self.h.rh_linearize(C.c_void_p(id(self.mvi)), self.solve, self.deriv1, C.c_long(B),
                            C.c_int(self.nq), C.c_int(self.nd), C.c_int(self.nk), C.c_int(self.nu),
                            C.c_int(self.nc), C.c_double(t1), C.c_double(dt), C.c_int(max_it),
                            _p(q1), _p(p1), _p(u1), _p(k2), _p(q2_hint), _p(lam_hint), _p(A),
                            _p(Bm) if nU else None, _p(it), _p(st))
        return dict(A=A, B=Bm, iters=it, status=st)


# ---- multi-process timing ---------------------------------------------------------------------
def _worker(args):
    name, kind, payload, reps = args
    h = Harness(name)
    t0 = time.perf_counter()
    units = 0
    for _ in range(reps):
        if kind == "rollouts":
            q, p, nsteps, dt = payload
            h.rollouts(q, p, nsteps, dt, dt)
            units += q.shape[0] * nsteps
        else:
            q1, p1, u1, k2, dt, lam = payload
            h.linearize(q1, p1, u1, k2, 0.0, dt, lam_hint=lam)
            units += q1.shape[0]
    return units, time.perf_counter() - t0


def time_parallel(name, kind, shards, reps=1, procs=None):
    """Run one shard per process; returns (units/s over all processes, processes used, wall s)."""
    procs = procs or len(shards)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(name, kind, s, reps) for s in shards])
    wall = time.perf_counter() - t0
    units = sum(r[0] for r in res)
    # throughput from the slowest worker's own loop time (excludes process start / system build)
    loop = max(r[1] for r in res)
    return units / loop, procs, wall


# ---- multi-process results (parity at scale: tests/test_gpu_parity_r2.py) ------------------------------
def _rollout_worker(args):
    name, q, p, nsteps, t0, dt = args
    return Harness(name).rollouts(q, p, nsteps, t0, dt)


def rollouts_parallel(name, q, p, nsteps, t0, dt, procs=None):
    """Harness.rollouts over all host cores (one reference MidpointVI per process, disjoint shards of the
    batch): final states, summed Newton iteration counts and status of every rollout."""
    procs = procs or len(os.sched_getaffinity(0))
    B = q.shape[0]
    cuts = np.linspace(0, B, procs + 1).astype(int)
    jobs = [(name, q[a:b], p[a:b], nsteps, t0, dt) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_rollout_worker, jobs)
    return {k: np.concatenate([r[k] for r in res]) for k in res[0]}
