"""TEST INFRASTRUCTURE: g++ build of the shared host/device math (tests/host_math_check.cc) so
the algebra of trep_b200/csrc/trepb_math.cuh can be checked against the reference's golden
vectors on a machine without a GPU.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from trep_b200 import desc as D

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libhostmath.so")
SRC = os.path.join(HERE, "host_math_check.cc")
DEPS = [SRC] + [os.path.join(ROOT, "trep_b200", "csrc", f)
                for f in ("trepb_math.cuh", "trepb_sys.h", "trepb_ws.h", "trepb_pack.h", "trepb_hd.h",
                          "trepb_d2.cuh", "trepb_d2jac.cuh", "trepb_kernels.cuh", "trepb_coop_math.cuh", "trepb_coop_sys.h", "trepb_coop.h")]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                               "-I/usr/local/cuda/include", "-x", "c++", SRC, "-o", SO])
    _lib = C.CDLL(SO)
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def step(desc, nsteps, t0, dt, q1, p1, u1=None, k2=None, q2_guess=None, lam_guess=None,
         tol=1e-10, maxit=200):
    lib = load()
    cd, keep = D.to_c(desc)
    q1, p1 = _c(q1), _c(p1)
    u1 = _c(np.zeros((nsteps, desc.nu)) if u1 is None else u1)
    k2 = _c(np.zeros((nsteps, desc.nk)) if k2 is None else k2)
    q2g = None if q2_guess is None else _c(q2_guess)
    lg = None if lam_guess is None else _c(lam_guess)
    q2, p2, lam = np.zeros(desc.nq), np.zeros(desc.nd), np.zeros(max(desc.nc, 1))
    it = C.c_int(0)
    lib.th_step.restype = C.c_int
    rc = lib.th_step(C.byref(cd), C.c_int(nsteps), C.c_double(t0), C.c_double(dt), C.c_double(tol),
                     C.c_int(maxit), _dp(q1), _dp(p1), _dp(u1), _dp(k2), _dp(q2g), _dp(lg),
                     _dp(q2), _dp(p2), _dp(lam), C.byref(it))
    return rc, q2, p2, lam[:desc.nc], it.value


def calc_p2(desc, dt, q0, q1):
    lib = load()
    cd, keep = D.to_c(desc)
    q0, q1 = _c(q0), _c(q1)
    p = np.zeros(desc.nd)
    rc = lib.th_calc_p2(C.byref(cd), C.c_double(dt), _dp(q0), _dp(q1), _dp(p))
    assert rc == 0
    return p


RAW = ["q2_dq1", "q2_dp1", "q2_du1", "q2_dk2", "p2_dq1", "p2_dp1", "p2_du1", "p2_dk2",
       "l1_dq1", "l1_dp1", "l1_du1", "l1_dk2"]


def raw_shapes(desc):
    nq, nd, nu, nk, nc = desc.nq, desc.nd, desc.nu, desc.nk, desc.nc
    wrt = {"dq1": nq, "dp1": nd, "du1": nu, "dk2": nk}
    return {n: (wrt[n[3:]], nc if n.startswith("l1") else nd) for n in RAW}


def linearize(desc, t1, t2, q1, p1, u1, k2, q2_guess=None, lam_guess=None, tol=1e-10, maxit=200):
    lib = load()
    cd, keep = D.to_c(desc)
    q1, p1, u1, k2 = _c(q1), _c(p1), _c(u1), _c(k2)
    q2g = None if q2_guess is None else _c(q2_guess)
    lg = None if lam_guess is None else _c(lam_guess)
    q2, p2, lam = np.zeros(desc.nq), np.zeros(desc.nd), np.zeros(max(desc.nc, 1))
    it = C.c_int(0)
    A = np.zeros((desc.nX, desc.nX))
    B = np.zeros((desc.nX, max(desc.nU, 1)))
    shapes = raw_shapes(desc)
    raw = {n: np.zeros(max(int(np.prod(shapes[n])), 1)) for n in RAW}
    ptrs = (C.POINTER(C.c_double) * 12)(*[_dp(raw[n]) for n in RAW])
    lib.th_linearize.restype = C.c_int
    rc = lib.th_linearize(C.byref(cd), C.c_double(t1), C.c_double(t2), C.c_double(tol), C.c_int(maxit),
                          _dp(q1), _dp(p1), _dp(u1), _dp(k2), _dp(q2g), _dp(lg), _dp(q2), _dp(p2),
                          _dp(lam), C.byref(it), _dp(A), _dp(B), ptrs)
    out = {n: raw[n][:int(np.prod(shapes[n]))].reshape(shapes[n]) for n in RAW}
    out.update(rc=rc, q2=q2, p2=p2, lambda1=lam[:desc.nc], iters=it.value, A=A,
               B=B[:, :desc.nU])
    return out


D2_KINDS = ["dq1dq1", "dq1dp1", "dq1du1", "dq1dk2", "dp1dp1", "dp1du1", "dp1dk2", "du1du1", "du1dk2", "dk2dk2"]
D2_WHICH = ["q2", "p2", "l1"]


def deriv2(desc, t1, t2, q1, p1, u1, k2, q2_guess=None, lam_guess=None, tol=1e-10, maxit=200, method="pair"):
    lib = load()
    cd, keep = D.to_c(desc)
    q1, p1, u1, k2 = _c(q1), _c(p1), _c(u1), _c(k2)
    q2g = None if q2_guess is None else _c(q2_guess)
    lg = None if lam_guess is None else _c(lam_guess)
    cnt = {"dq1": desc.nq, "dp1": desc.nd, "du1": desc.nu, "dk2": desc.nk}
    out, ptrs = {}, []
    for w in D2_WHICH:
        for kd in D2_KINDS:
            sh = (cnt[kd[:3]], cnt[kd[3:]], desc.nc if w == "l1" else desc.nd)
            buf = np.zeros(max(int(np.prod(sh)), 1))
            out[w + "_" + kd] = (buf, sh)
            ptrs.append(_dp(buf))
    arr = (C.POINTER(C.c_double) * 30)(*ptrs)
    fn = lib.th_deriv2 if method == "pair" else lib.th_deriv2_jac
    fn.restype = C.c_int
    rc = fn(C.byref(cd), C.c_double(t1), C.c_double(t2), C.c_double(tol), C.c_int(maxit),
                       _dp(q1), _dp(p1), _dp(u1), _dp(k2), _dp(q2g), _dp(lg), arr)
    res = {n: b[:int(np.prod(sh))].reshape(sh) for n, (b, sh) in out.items()}
    res["rc"] = rc
    return res


def coop_info(desc):
    lib = load()
    cd, keep = D.to_c(desc)
    out = (C.c_int * 8)()
    rc = lib.th_coop_info(C.byref(cd), out)
    if rc:
        return None
    return dict(nl=out[0], nlevels=out[1], npairs=out[2], npoints=out[3], ws_doubles=out[4], blob_bytes=out[5],
                ws_doubles_static=out[6])


def coop_linearize(desc, t1, t2, q1, p1, u1, k2, q2_guess=None, lam_guess=None, tol=1e-10, maxit=200,
                   nsteps=1, derivs=True, static_dims=False):
    """Same call as linearize() through the team-cooperative math (one-lane host team)."""
    lib = load()
    cd, keep = D.to_c(desc)
    q1, p1, u1, k2 = _c(q1), _c(p1), _c(u1), _c(k2)
    q2g = None if q2_guess is None else _c(q2_guess)
    lg = None if lam_guess is None else _c(lam_guess)
    q2, p2, lam = np.zeros(desc.nq), np.zeros(desc.nd), np.zeros(max(desc.nc, 1))
    it = C.c_int(0)
    A = np.full((desc.nX, desc.nX), 1e300)          # poisoned: every entry of A and B must be written
    B = np.full((desc.nX, max(desc.nU, 1)), 1e300)
    shapes = raw_shapes(desc)
    raw = {n: np.zeros(max(int(np.prod(shapes[n])), 1)) for n in RAW}
    ptrs = (C.POINTER(C.c_double) * 12)(*[_dp(raw[n]) for n in RAW])
    aux = np.zeros(4 * (desc.nd + desc.nc + 2) ** 2)
    lib.th_coop_linearize.restype = C.c_int
    rc = lib.th_coop_linearize(C.byref(cd), C.c_int(int(static_dims)), C.c_int(nsteps), C.c_double(t1), C.c_double(t2 - t1), C.c_double(tol),
                               C.c_int(maxit), _dp(q1), _dp(p1), _dp(u1), _dp(k2), _dp(q2g), _dp(lg), _dp(q2),
                               _dp(p2), _dp(lam), C.byref(it), _dp(A), _dp(B), ptrs if derivs else None, _dp(aux))
    out = {n: raw[n][:int(np.prod(shapes[n]))].reshape(shapes[n]) for n in RAW}
    out.update(rc=rc, q2=q2, p2=p2, lambda1=lam[:desc.nc], iters=it.value, A=A, B=B[:, :desc.nU], aux=aux)
    return out


def coop_calc_p2(desc, dt, q0, q1):
    lib = load()
    cd, keep = D.to_c(desc)
    q0, q1 = _c(q0), _c(q1)
    p = np.zeros(desc.nd)
    rc = lib.th_coop_calc_p2(C.byref(cd), C.c_double(dt), _dp(q0), _dp(q1), _dp(p))
    assert rc == 0
    return p
