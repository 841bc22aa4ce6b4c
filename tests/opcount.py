"""TEST INFRASTRUCTURE: operation counts of the thread-per-instance math (tests/opcount.cc: trepb_math.cuh compiled
with `double` replaced by a counting wrapper).  Used by tests/test_opcount.py and tools/count_ops.py (which writes
profiles/opcounts.json for bench.py's roofline notes).  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from trep_b200 import desc as D

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libopcount.so")
SRC = os.path.join(HERE, "opcount.cc")
DEPS = [SRC] + [os.path.join(ROOT, "trep_b200", "csrc", f) for f in ("trepb_math.cuh", "trepb_sys.h", "trepb_ws.h", "trepb_pack.h")]
NAMES = ("add", "mul", "div", "sqrt", "fma", "sincos", "cmp")
_lib = None


def load():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                                   "-I/usr/local/cuda/include", SRC, "-o", SO])
        _lib = C.CDLL(SO)
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def flops(counts):
    """add, sub, mul, div, sqrt = 1 each, fma = 2; sin / cos evaluations are reported separately."""
    return counts["add"] + counts["mul"] + counts["div"] + counts["sqrt"] + 2 * counts["fma"]


def step(desc, nsteps, t0, dt, q1, p1, u1=None, k2=None, lam_guess=None, tol=1e-10, maxit=200):
    lib = load()
    cd, keep = D.to_c(desc)
    q1, p1 = _c(q1), _c(p1)
    u1 = _c(np.zeros((nsteps, desc.nu)) if u1 is None else u1)
    k2 = _c(np.zeros((nsteps, desc.nk)) if k2 is None else k2)
    lg = None if lam_guess is None else _c(lam_guess)
    q2, p2 = np.zeros(desc.nq), np.zeros(desc.nd)
    cnt = (C.c_ulonglong * 7)()
    lib.oc_step.restype = C.c_int
    it = lib.oc_step(C.byref(cd), C.c_int(nsteps), C.c_double(t0), C.c_double(dt), C.c_double(tol), C.c_int(maxit),
                     _dp(q1), _dp(p1), _dp(u1), _dp(k2), _dp(lg), _dp(q2), _dp(p2), cnt)
    return dict(iters=it, q2=q2, p2=p2, counts=dict(zip(NAMES, [int(x) for x in cnt])))


def linearize(desc, t1, t2, q1, p1, u1, k2, q2_guess=None, lam_guess=None, tol=1e-10, maxit=200):
    lib = load()
    cd, keep = D.to_c(desc)
    q1, p1, u1, k2 = _c(q1), _c(p1), _c(u1), _c(k2)
    q2g = None if q2_guess is None else _c(q2_guess)
    lg = None if lam_guess is None else _c(lam_guess)
    A = np.zeros((desc.nX, desc.nX)); B = np.zeros((desc.nX, max(desc.nU, 1)))
    cnt = (C.c_ulonglong * 7)()
    lib.oc_linearize.restype = C.c_int
    it = lib.oc_linearize(C.byref(cd), C.c_double(t1), C.c_double(t2), C.c_double(tol), C.c_int(maxit), _dp(q1), _dp(p1),
                          _dp(u1), _dp(k2), _dp(q2g), _dp(lg), _dp(A), _dp(B), cnt)
    return dict(iters=it, A=A, B=B[:, :desc.nU], counts=dict(zip(NAMES, [int(x) for x in cnt])))
