"""ctypes binding of libtrepb.so (the C ABI declared in include/trepb.h).

There is no fallback: if the library has not been built (``python -m trep_b200.build``) the
import of this module raises, and every compute call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import desc as D

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("TREPB_LIBPATH") or os.path.join(HERE, "libtrepb.so")

if not os.path.exists(LIBPATH):
    raise ImportError("trep_b200/libtrepb.so is missing: build it with `python -m trep_b200.build` "
                      "(there is no CPU fallback)")
_lib = C.CDLL(LIBPATH)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class StepArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("nsteps", C.c_int32), ("max_iterations", C.c_int32),
                ("t0", C.c_double), ("dt", C.c_double), ("tolerance", C.c_double),
                ("q1", C.c_void_p), ("p1", C.c_void_p), ("u1", C.c_void_p), ("k2", C.c_void_p),
                ("q2_guess", C.c_void_p), ("lambda_guess", C.c_void_p),
                ("q2", C.c_void_p), ("p2", C.c_void_p), ("lambda1", C.c_void_p),
                ("iters", C.c_void_p), ("status", C.c_void_p),
                ("sample_every", C.c_int32), ("_pad", C.c_int32),
                ("traj_q", C.c_void_p), ("traj_p", C.c_void_p), ("times", C.c_void_p)]


class ProjectArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("nsteps", C.c_int32), ("max_iterations", C.c_int32),
                ("t0", C.c_double), ("dt", C.c_double), ("tolerance", C.c_double),
                ("bX", C.c_void_p), ("bU", C.c_void_p), ("Kfb", C.c_void_p),
                ("k_per_instance", C.c_int32), ("use_hint", C.c_int32),
                ("X", C.c_void_p), ("U", C.c_void_p),
                ("iters", C.c_void_p), ("status", C.c_void_p), ("fail_step", C.c_void_p), ("times", C.c_void_p)]


class LqrArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("nsteps", C.c_int32), ("nX", C.c_int32), ("nU", C.c_int32),
                ("q_per_step", C.c_int32), ("r_per_step", C.c_int32), ("_pad", C.c_int32),
                ("A", C.c_void_p), ("B", C.c_void_p), ("Q", C.c_void_p), ("R", C.c_void_p),
                ("Kfb", C.c_void_p), ("P0", C.c_void_p), ("status", C.c_void_p)]


class LqArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("nsteps", C.c_int32), ("nX", C.c_int32), ("nU", C.c_int32),
                ("cost_per_rollout", C.c_int32),
                ("A", C.c_void_p), ("B", C.c_void_p), ("Q", C.c_void_p), ("S", C.c_void_p), ("R", C.c_void_p),
                ("q", C.c_void_p), ("r", C.c_void_p), ("Kfb", C.c_void_p), ("C", C.c_void_p), ("P0", C.c_void_p),
                ("b0", C.c_void_p), ("status", C.c_void_p)]


RAW = ["q2_dq1", "q2_dp1", "q2_du1", "q2_dk2", "p2_dq1", "p2_dp1", "p2_du1", "p2_dk2",
       "l1_dq1", "l1_dp1", "l1_du1", "l1_dk2"]


class LinArgs(C.Structure):
    _fields_ = ([("batch", C.c_int64), ("max_iterations", C.c_int32), ("traj_len", C.c_int32),
                 ("tolerance", C.c_double), ("t1", C.c_void_p), ("t2", C.c_void_p),
                 ("t1_scalar", C.c_double), ("dt_scalar", C.c_double),
                 ("q1", C.c_void_p), ("p1", C.c_void_p), ("u1", C.c_void_p), ("k2", C.c_void_p),
                 ("q2_guess", C.c_void_p), ("lambda_guess", C.c_void_p),
                 ("q2", C.c_void_p), ("p2", C.c_void_p), ("lambda1", C.c_void_p),
                 ("iters", C.c_void_p), ("status", C.c_void_p), ("A", C.c_void_p), ("B", C.c_void_p)]
                + [(n, C.c_void_p) for n in RAW])


D2_KINDS = ["dq1dq1", "dq1dp1", "dq1du1", "dq1dk2", "dp1dp1", "dp1du1", "dp1dk2", "du1du1", "du1dk2", "dk2dk2"]
D2_WHICH = ["q2", "p2", "l1"]


class D2Args(C.Structure):
    _fields_ = [("lin", LinArgs), ("d2", C.c_void_p * 30), ("z", C.c_void_p), ("fdxdx", C.c_void_p),
                ("fdxdu", C.c_void_p), ("fdudu", C.c_void_p)]


_lib.trepb_last_error.restype = C.c_char_p
_lib.trepb_system_kernel_name.restype = C.c_char_p
_lib.trepb_specialized_name.restype = C.c_char_p
_lib.trepb_desc_hash.restype = C.c_uint64
_lib.trepb_struct_hash.restype = C.c_uint64
_lib.trepb_system_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
_lib.trepb_system_destroy.argtypes = [C.c_void_p]
_lib.trepb_system_destroy.restype = None
_lib.trepb_system_is_specialized.argtypes = [C.c_void_p]
_lib.trepb_system_is_cooperative.argtypes = [C.c_void_p]
_lib.trepb_system_kernel_name.argtypes = [C.c_void_p]
_lib.trepb_step_batch.argtypes = [C.c_void_p, C.POINTER(StepArgs)]
_lib.trepb_step_batch_dev.argtypes = [C.c_void_p, C.POINTER(StepArgs), C.c_void_p]
_lib.trepb_project_batch.argtypes = [C.c_void_p, C.POINTER(ProjectArgs)]
_lib.trepb_project_batch_dev.argtypes = [C.c_void_p, C.POINTER(ProjectArgs), C.c_void_p]
_lib.trepb_lqr_batch.argtypes = [C.c_int, C.POINTER(LqrArgs)]
_lib.trepb_lqr_batch_dev.argtypes = [C.c_int, C.POINTER(LqrArgs), C.c_void_p]
_lib.trepb_lq_batch.argtypes = [C.c_int, C.POINTER(LqArgs)]
_lib.trepb_lq_batch_dev.argtypes = [C.c_int, C.POINTER(LqArgs), C.c_void_p]
_lib.trepb_lqr_last_kernel_ms.argtypes = [C.c_int, C.POINTER(C.c_float)]
_lib.trepb_linearize_batch.argtypes = [C.c_void_p, C.POINTER(LinArgs)]
_lib.trepb_linearize_batch_dev.argtypes = [C.c_void_p, C.POINTER(LinArgs), C.c_void_p]
_lib.trepb_deriv2_batch.argtypes = [C.c_void_p, C.POINTER(D2Args)]
_lib.trepb_deriv2_batch_dev.argtypes = [C.c_void_p, C.POINTER(D2Args), C.c_void_p]
_lib.trepb_calc_p2_batch.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
_lib.trepb_calc_p2_batch_dev.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
_lib.trepb_calc_f_batch.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double] + [C.c_void_p] * 6
_lib.trepb_calc_f_batch_dev.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double] + [C.c_void_p] * 7
_lib.trepb_discrete_fm2_batch.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double] + [C.c_void_p] * 4
_lib.trepb_discrete_fm2_batch_dev.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double] + [C.c_void_p] * 5
_lib.trepb_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
_lib.trepb_kernel_info.argtypes = [C.c_void_p, C.c_int] + [_ip] * 5
_lib.trepb_malloc.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
_lib.trepb_free.argtypes = [C.c_int, C.c_void_p]
_lib.trepb_host_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p)]
_lib.trepb_host_free.argtypes = [C.c_void_p]
_lib.trepb_memcpy_h2d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
_lib.trepb_memcpy_d2h.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
_lib.trepb_memset.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int64]
_lib.trepb_memcpy_d2d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
_lib.trepb_comm_available.argtypes = [C.POINTER(C.c_int)]
_lib.trepb_comm_unique_id.argtypes = [C.c_char_p]
_lib.trepb_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
_lib.trepb_comm_destroy.argtypes = [C.c_void_p]
_lib.trepb_comm_destroy.restype = None
_lib.trepb_comm_rank.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
_lib.trepb_comm_allgather_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
_lib.trepb_comm_gather_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
_lib.trepb_ipc_export.argtypes = [C.c_int, C.c_void_p, C.c_char_p]
_lib.trepb_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
_lib.trepb_ipc_close.argtypes = [C.c_int, C.c_void_p]
_lib.trepb_event_pair_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
_lib.trepb_event_record.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
_lib.trepb_event_elapsed_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
_lib.trepb_event_pair_destroy.argtypes = [C.c_void_p]
_lib.trepb_event_pair_destroy.restype = None
_lib.trepb_measure_fp64_peak.argtypes = [C.c_int, _dp]
_lib.trepb_sincos_batch.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]

EXPORTS = [
    "trepb_abi_version", "trepb_last_error", "trepb_system_create", "trepb_system_destroy",
    "trepb_system_dims", "trepb_system_is_specialized", "trepb_system_is_cooperative", "trepb_system_kernel_name",
    "trepb_kernel_info", "trepb_validate", "trepb_codegen", "trepb_codegen_literal", "trepb_desc_hash", "trepb_struct_hash", "trepb_coop_dims",
    "trepb_num_specialized", "trepb_specialized_name", "trepb_step_batch", "trepb_step_batch_dev",
    "trepb_calc_p2_batch", "trepb_calc_p2_batch_dev", "trepb_calc_f_batch", "trepb_calc_f_batch_dev",
    "trepb_discrete_fm2_batch", "trepb_discrete_fm2_batch_dev", "trepb_project_batch", "trepb_project_batch_dev", "trepb_lqr_batch", "trepb_lqr_batch_dev", "trepb_lq_batch", "trepb_lq_batch_dev", "trepb_lqr_last_kernel_ms", "trepb_load_plugin", "trepb_linearize_batch",
    "trepb_linearize_batch_dev", "trepb_deriv2_batch", "trepb_deriv2_batch_dev", "trepb_device_count", "trepb_malloc", "trepb_free",
    "trepb_host_alloc", "trepb_host_free", "trepb_memset", "trepb_memcpy_h2d", "trepb_memcpy_d2h", "trepb_memcpy_d2d",
    "trepb_comm_available", "trepb_comm_unique_id", "trepb_comm_create", "trepb_comm_destroy", "trepb_comm_rank",
    "trepb_comm_allgather_dev", "trepb_comm_gather_dev", "trepb_ipc_export", "trepb_ipc_open", "trepb_ipc_close",
    "trepb_event_pair_create", "trepb_event_record", "trepb_event_elapsed_ms", "trepb_event_pair_destroy",
    "trepb_synchronize", "trepb_last_kernel_ms", "trepb_measure_fp64_peak", "trepb_sincos_batch",
]


class TrepbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("trepb error %d: %s" % (code, msg))
        self.code = code


def _check(rc):
    if rc != 0:
        raise TrepbError(rc, (_lib.trepb_last_error() or b"").decode())


def raw():
    """The ctypes library object (for symbol checks)."""
    return _lib


def device_count():
    n = C.c_int(0)
    rc = _lib.trepb_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def coop_dims(desc):
    """Shape the compile-time-size cooperative kernels are instantiated for: (nd, nk, nu, nc, links, points,
    chain pairs, levels, constrained dynamic configs, constrained configs, extras), or None when the cooperative
    kernels do not apply; extras = 1 for a system with LinearSprings / LinearDampers / spline springs / wrenches
    (their counts stay run-time data of the instantiation)."""
    cd, keep = D.to_c(desc)
    out = (C.c_int32 * 11)()
    if _lib.trepb_coop_dims(C.byref(cd), out) != 0:
        return None
    return tuple(out)


def lqr_raw(on_device, device, batch, nsteps, nX, nU, A, B, Q, R, Kfb, status, P0=None, q_per_step=False,
            r_per_step=False, stream=None):
    a = LqrArgs(batch=batch, nsteps=nsteps, nX=nX, nU=nU, q_per_step=1 if q_per_step else 0,
                r_per_step=1 if r_per_step else 0, A=_ptr(A), B=_ptr(B), Q=_ptr(Q), R=_ptr(R), Kfb=_ptr(Kfb),
                P0=_ptr(P0), status=_ptr(status))
    if on_device:
        _check(_lib.trepb_lqr_batch_dev(device, C.byref(a), stream))
    else:
        _check(_lib.trepb_lqr_batch(device, C.byref(a)))


def load_plugin(path):
    """Load a plug-in of specialised kernels (trep_b200.build.build_plugin); returns the number of kernel sets it added."""
    _lib.trepb_load_plugin.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    n = C.c_int(0)
    _check(_lib.trepb_load_plugin(os.fspath(path).encode(), C.byref(n)))
    return int(n.value)


def specialize(desc, kind="auto", ext=False, verbose=False):
    """Builds (once per structure: the library is named by trepb_struct_hash and rebuilt only when older than the
    headers) and loads a plug-in of kernels compiled for `desc`'s structure, so that System(desc) - and every other
    description with the same structure, whatever its masses, lengths and gains - runs on them instead of the
    table-driven kernels.  kind / ext as trep_b200.build.build_plugin.  Needs nvcc (a minute per structure); raises
    RuntimeError without it: there is no silent fallback.  Returns the number of kernel sets added (0 if this
    structure's plug-in was already loaded)."""
    import shutil
    from . import build
    if kind == "auto":
        kind = "thread" if desc.nd + desc.nk <= 6 else "coop"
    name = "s%016x_%s%s" % (struct_hash(desc), kind, "_ext" if ext else "")
    if name in _specialized:
        return 0
    if shutil.which(build.NVCC) is None:
        raise RuntimeError("trep_b200.lib.specialize needs nvcc (%s) to build the plug-in" % build.NVCC)
    n = load_plugin(build.build_plugin(desc, name, verbose=verbose, kind=kind, ext=ext))
    _specialized.add(name)
    return n


_specialized = set()


def lqr_last_kernel_ms(device=0):
    """Device time (CUDA events) of the last Riccati kernel launched on `device`."""
    ms = C.c_float()
    _check(_lib.trepb_lqr_last_kernel_ms(device, C.byref(ms)))
    return float(ms.value)


def solve_tv_lqr(A, B, Q, R, device=0):
    """Batched discopt.dlqr.solve_tv_lqr (trep/discopt/dlqr.py:9-38) on the GPU.
    A [R,K,nX,nX] or [K,nX,nX]; B likewise; Q [nX,nX] or [K+1,nX,nX]; R [nU,nU] or [K,nU,nU].
    Returns (K [..,K,nU,nX], P0 [..,nX,nX])."""
    A = np.ascontiguousarray(np.asarray(A, float)); B = np.ascontiguousarray(np.asarray(B, float))
    single = A.ndim == 3
    if single:
        A, B = A[None], B[None]
    Rn, K, nX = A.shape[0], A.shape[1], A.shape[2]
    nU = B.shape[3]
    Q = np.ascontiguousarray(np.asarray(Q, float)); R = np.ascontiguousarray(np.asarray(R, float))
    Kfb = np.zeros((Rn, K, nU, nX)); P0 = np.zeros((Rn, nX, nX)); status = np.zeros(Rn, np.int32)
    lqr_raw(False, device, Rn, K, nX, nU, A, B, Q, R, Kfb, status, P0=P0, q_per_step=Q.ndim == 3, r_per_step=R.ndim == 3)
    if np.any(status != 0):
        raise TrepbError(0, "singular gamma in the Riccati sweep of rollout(s) %s" % np.flatnonzero(status != 0)[:8])
    return (Kfb[0], P0[0]) if single else (Kfb, P0)


def lq_raw(on_device, device, batch, nsteps, nX, nU, A, B, Q, S, R, q, r, Kfb, Cff, status, P0=None, b0=None,
           cost_per_rollout=False, stream=None):
    a = LqArgs(batch=batch, nsteps=nsteps, nX=nX, nU=nU, cost_per_rollout=1 if cost_per_rollout else 0,
               A=_ptr(A), B=_ptr(B), Q=_ptr(Q), S=_ptr(S), R=_ptr(R), q=_ptr(q), r=_ptr(r), Kfb=_ptr(Kfb),
               C=_ptr(Cff), P0=_ptr(P0), b0=_ptr(b0), status=_ptr(status))
    if on_device:
        _check(_lib.trepb_lq_batch_dev(device, C.byref(a), stream))
    else:
        _check(_lib.trepb_lq_batch(device, C.byref(a)))


def solve_tv_lq(A, B, q, r, Q, S, R, device=0):
    """Batched discopt.dlqr.solve_tv_lq (trep/discopt/dlqr.py:41-81) on the GPU.
    One rollout: A [K,nX,nX], B [K,nX,nU], q [K+1,nX], r [K,nU], Q [K+1,nX,nX], S [K,nX,nU] or None,
    R [K,nU,nU]; a batch: one more leading axis on A and B, and either the same on every cost array
    (one cost per rollout) or none (costs shared).  Returns (K, C, P0, b0)."""
    f = lambda x: np.ascontiguousarray(np.asarray(x, float))
    A, B, q, r, Q, R = f(A), f(B), f(q), f(r), f(Q), f(R)
    S = None if S is None else f(S)
    single = A.ndim == 3
    if single:
        A, B = A[None], B[None]
    Rn, K, nX = A.shape[0], A.shape[1], A.shape[2]
    nU = B.shape[3]
    per = Q.ndim == 4
    Kfb = np.zeros((Rn, K, nU, nX)); Cff = np.zeros((Rn, K, nU)); P0 = np.zeros((Rn, nX, nX)); b0 = np.zeros((Rn, nX))
    status = np.zeros(Rn, np.int32)
    lq_raw(False, device, Rn, K, nX, nU, A, B, Q, S, R, q, r, Kfb, Cff, status, P0=P0, b0=b0, cost_per_rollout=per)
    if np.any(status != 0):
        raise TrepbError(0, "singular gamma in the Riccati sweep of rollout(s) %s" % np.flatnonzero(status != 0)[:8])
    return (Kfb[0], Cff[0], P0[0], b0[0]) if single else (Kfb, Cff, P0, b0)


def specialized_names():
    return [_lib.trepb_specialized_name(i).decode() for i in range(_lib.trepb_num_specialized())]


def desc_hash(desc):
    cd, keep = D.to_c(desc)
    return int(_lib.trepb_desc_hash(C.byref(cd)))


def struct_hash(desc):
    """Hash of the structure only (what a specialised kernel is compiled for); see include/trepb.h."""
    cd, keep = D.to_c(desc)
    return int(_lib.trepb_struct_hash(C.byref(cd)))


def validate(desc):
    cd, keep = D.to_c(desc)
    _check(_lib.trepb_validate(C.byref(cd)))


def measure_fp64_peak(device=0):
    v = C.c_double(0)
    _check(_lib.trepb_measure_fp64_peak(device, C.byref(v)))
    return v.value


def sincos(x, device=0):
    """(sin, cos) of an array as the kernels compute them (diagnostic)."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    s, c = np.empty_like(x), np.empty_like(x)
    _check(_lib.trepb_sincos_batch(device, x.size, x.ctypes.data, s.ctypes.data, c.ctypes.data))
    return s, c


def synchronize(device=0):
    _check(_lib.trepb_synchronize(device))


class EventTimer:
    """CUDA-event pair on a stream (trepb_event_*): `with EventTimer(dev) as t: ...launches...; t.ms`."""

    def __init__(self, device=0, stream=None):
        self.device, self.stream = device, stream
        self._h = C.c_void_p()
        _check(_lib.trepb_event_pair_create(device, C.byref(self._h)))
        self.ms = None

    def __enter__(self):
        _check(_lib.trepb_event_record(self._h, 0, self.stream))
        return self

    def __exit__(self, *exc):
        _check(_lib.trepb_event_record(self._h, 1, self.stream))
        v = C.c_float(0)
        _check(_lib.trepb_event_elapsed_ms(self._h, C.byref(v)))
        self.ms = v.value
        return False

    def __del__(self):
        try:
            if self._h:
                _lib.trepb_event_pair_destroy(self._h)
                self._h = None
        except Exception:
            pass


class DeviceBuffer:
    """HBM buffer owned through the C ABI (so that a host without torch can drive the library)."""

    def __init__(self, device, shape, dtype=np.float64):
        self.device, self.shape, self.dtype = device, tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        _check(_lib.trepb_malloc(device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def data_ptr(self):
        return self.ptr

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.nbytes == self.nbytes, (host.shape, self.shape)
        _check(_lib.trepb_memcpy_h2d(self.device, self.ptr, host.ctypes.data, self.nbytes))
        return self

    def download(self, out=None):
        if out is None:
            out = np.empty(self.shape, self.dtype)
        _check(_lib.trepb_memcpy_d2h(self.device, out.ctypes.data, self.ptr, self.nbytes))
        return out

    def zero(self):
        _check(_lib.trepb_memset(self.device, self.ptr, 0, self.nbytes))
        return self

    def free(self):
        if self.ptr:
            _lib.trepb_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by pinned host memory (cudaHostAlloc) for the end-to-end path."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = C.c_void_p()
    _check(_lib.trepb_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
    return arr


def _ptr(x):
    """Address of a numpy array / DeviceBuffer / torch tensor, or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()


class System:
    """Handle of a flattened system resident on one GPU (trepb_system)."""

    def __init__(self, desc: D.SystemDesc, device=0, specialize=True, cooperative=None, d2_pairwise=False, literal=True,
                 coop_one_warp=False, coop_two_warps=False):
        """specialize=False: skip the ahead-of-time specialised kernels.  cooperative: None = let
        the library choose between one thread and one warp per instance for a table-driven system,
        False = always one thread, True = always the cooperative kernels (their compile-time-size
        flavour when one was built for this shape, unless specialize=False).  d2_pairwise: second
        derivatives of a table-driven system by one hyper-dual residual evaluation per parameter
        pair instead of one dual evaluation of the Jacobian tables per parameter.  coop_one_warp /
        coop_two_warps: for shapes whose cooperative kernels were built in several flavours, the one-warp
        flavour with the whole workspace in shared memory / the two-warps-per-instance flavour instead of
        the default (TREPB_FLAG_COOP_ONE_WARP, TREPB_FLAG_COOP_TWO_WARPS)."""
        self.desc = desc
        self.device = device
        cd, self._keep = D.to_c(desc)
        h = C.c_void_p()
        flags = 0 if specialize else 1
        if cooperative is False:
            flags |= 2
        elif cooperative is True:
            flags |= 4
        if d2_pairwise:
            flags |= 8
        if coop_one_warp:
            flags |= 32
        if coop_two_warps:
            flags |= 64
        if not literal:
            flags |= 16     # skip an all-literal instantiation: the run-time-parameter kernel of the structure
        _check(_lib.trepb_system_create(C.byref(cd), device, flags, C.byref(h)))
        self._h = h
        self.nq, self.nd, self.nk, self.nu, self.nc = desc.nq, desc.nd, desc.nk, desc.nu, desc.nc
        self.nX, self.nU = desc.nX, desc.nU

    @property
    def specialized(self):
        return bool(_lib.trepb_system_is_specialized(self._h))

    @property
    def cooperative(self):
        return bool(_lib.trepb_system_is_cooperative(self._h))

    @property
    def kernel_name(self):
        return _lib.trepb_system_kernel_name(self._h).decode()

    def kernel_info(self, which):
        v = [C.c_int32(0) for _ in range(5)]
        _check(_lib.trepb_kernel_info(self._h, which, *[C.byref(x) for x in v]))
        return dict(zip(["regs", "local_bytes", "blocks_per_sm", "block", "smem_bytes"], [x.value for x in v]))

    def last_kernel_ms(self):
        ms = C.c_float(0)
        _check(_lib.trepb_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def close(self):
        if getattr(self, "_h", None):
            _lib.trepb_system_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- raw calls: every array argument is a pointer-like (numpy => host entry point,
    #      anything with data_ptr() => device entry point); the caller owns shapes/dtypes ----------
    def step_raw(self, on_device, batch, nsteps, t0, dt, q1, p1, u1, k2, q2_guess, lambda_guess,
                 q2, p2, lambda1, iters, status, tolerance=1e-10, max_iterations=200,
                 sample_every=0, traj_q=None, traj_p=None, stream=None, times=None):
        a = StepArgs(times=_ptr(times), batch=batch, nsteps=nsteps, max_iterations=max_iterations, t0=t0, dt=dt,
                     tolerance=tolerance, q1=_ptr(q1), p1=_ptr(p1), u1=_ptr(u1), k2=_ptr(k2),
                     q2_guess=_ptr(q2_guess), lambda_guess=_ptr(lambda_guess), q2=_ptr(q2),
                     p2=_ptr(p2), lambda1=_ptr(lambda1), iters=_ptr(iters), status=_ptr(status),
                     sample_every=sample_every, traj_q=_ptr(traj_q), traj_p=_ptr(traj_p))
        if on_device:
            _check(_lib.trepb_step_batch_dev(self._h, C.byref(a), stream))
        else:
            _check(_lib.trepb_step_batch(self._h, C.byref(a)))

    @staticmethod
    def _lin_args(batch, q1, p1, u1, k2, status, t1=None, t2=None, t1_scalar=0.0, dt_scalar=0.0,
                  q2_guess=None, lambda_guess=None, q2=None, p2=None, lambda1=None, iters=None,
                  A=None, B=None, raw=None, tolerance=1e-10, max_iterations=200, traj_len=0):
        a = LinArgs(traj_len=traj_len, batch=batch, max_iterations=max_iterations, tolerance=tolerance, t1=_ptr(t1),
                    t2=_ptr(t2), t1_scalar=t1_scalar, dt_scalar=dt_scalar, q1=_ptr(q1), p1=_ptr(p1),
                    u1=_ptr(u1), k2=_ptr(k2), q2_guess=_ptr(q2_guess),
                    lambda_guess=_ptr(lambda_guess), q2=_ptr(q2), p2=_ptr(p2),
                    lambda1=_ptr(lambda1), iters=_ptr(iters), status=_ptr(status), A=_ptr(A), B=_ptr(B))
        for n, v in (raw or {}).items():
            setattr(a, n, _ptr(v))
        return a

    def linearize_raw(self, on_device, batch, q1, p1, u1, k2, status, stream=None, **kw):
        a = self._lin_args(batch, q1, p1, u1, k2, status, **kw)
        if on_device:
            _check(_lib.trepb_linearize_batch_dev(self._h, C.byref(a), stream))
        else:
            _check(_lib.trepb_linearize_batch(self._h, C.byref(a)))

    def deriv2_raw(self, on_device, batch, q1, p1, u1, k2, status, d2, stream=None, z=None,
                   fdxdx=None, fdxdu=None, fdudu=None, **kw):
        """d2: {"q2_dq1dq1": array, ...} any subset of D2_WHICH x D2_KINDS; z + fdxdx/fdxdu/fdudu:
        the z-contracted forms of DSystem.fdxdx(z) / fdxdu(z) / fdudu(z)."""
        a = D2Args()
        a.lin = self._lin_args(batch, q1, p1, u1, k2, status, **kw)
        a.z, a.fdxdx, a.fdxdu, a.fdudu = _ptr(z), _ptr(fdxdx), _ptr(fdxdu), _ptr(fdudu)
        for n, v in d2.items():
            w, kd = n.split("_")
            a.d2[10 * D2_WHICH.index(w) + D2_KINDS.index(kd)] = _ptr(v)
        if on_device:
            _check(_lib.trepb_deriv2_batch_dev(self._h, C.byref(a), stream))
        else:
            _check(_lib.trepb_deriv2_batch(self._h, C.byref(a)))

    def project_raw(self, on_device, batch, nsteps, t0, dt, bX, bU, Kfb, X, U, status, k_per_instance=False,
                    use_hint=True, iters=None, fail_step=None, tolerance=1e-10, max_iterations=200, stream=None,
                    times=None):
        a = ProjectArgs(times=_ptr(times), batch=batch, nsteps=nsteps, max_iterations=max_iterations, t0=t0, dt=dt, tolerance=tolerance,
                        bX=_ptr(bX), bU=_ptr(bU), Kfb=_ptr(Kfb), k_per_instance=1 if k_per_instance else 0,
                        use_hint=1 if use_hint else 0, X=_ptr(X), U=_ptr(U), iters=_ptr(iters), status=_ptr(status),
                        fail_step=_ptr(fail_step))
        if on_device:
            _check(_lib.trepb_project_batch_dev(self._h, C.byref(a), stream))
        else:
            _check(_lib.trepb_project_batch(self._h, C.byref(a)))

    def calc_p2_raw(self, on_device, batch, dt, q0, q1, p, stream=None):
        if on_device:
            _check(_lib.trepb_calc_p2_batch_dev(self._h, batch, dt, _ptr(q0), _ptr(q1), _ptr(p), stream))
        else:
            _check(_lib.trepb_calc_p2_batch(self._h, batch, dt, _ptr(q0), _ptr(q1), _ptr(p)))

    def calc_f_raw(self, on_device, batch, t1, t2, q1, q2, p1, u1, lambda1, f, stream=None):
        if on_device:
            _check(_lib.trepb_calc_f_batch_dev(self._h, batch, t1, t2, _ptr(q1), _ptr(q2), _ptr(p1), _ptr(u1),
                                               _ptr(lambda1), _ptr(f), stream))
        else:
            _check(_lib.trepb_calc_f_batch(self._h, batch, t1, t2, _ptr(q1), _ptr(q2), _ptr(p1), _ptr(u1),
                                           _ptr(lambda1), _ptr(f)))

    def discrete_fm2_raw(self, on_device, batch, t1, t2, q1, q2, u1, fm2, stream=None):
        if on_device:
            _check(_lib.trepb_discrete_fm2_batch_dev(self._h, batch, t1, t2, _ptr(q1), _ptr(q2), _ptr(u1), _ptr(fm2), stream))
        else:
            _check(_lib.trepb_discrete_fm2_batch(self._h, batch, t1, t2, _ptr(q1), _ptr(q2), _ptr(u1), _ptr(fm2)))

    # ---- numpy convenience (host entry points) ---------------------------------------------------
    def _f(self, x, shape):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        return x.reshape(shape)

    def calc_p2(self, dt, q0, q1):
        q0 = np.atleast_2d(np.asarray(q0, float))
        B = q0.shape[0]
        q0, q1 = self._f(q0, (B, self.nq)), self._f(q1, (B, self.nq))
        p = np.empty((B, self.nd))
        self.calc_p2_raw(False, B, float(dt), q0, q1, p)
        return p

    def calc_f(self, t1, t2, q1, q2, p1, u1=None, lambda1=None):
        """Residual of the DEL equation [B][nd+nc] (_MidpointVI._calc_f, midpointvi.c:533-575)."""
        q1 = np.atleast_2d(np.asarray(q1, float))
        B = q1.shape[0]
        q1, q2, p1 = self._f(q1, (B, self.nq)), self._f(q2, (B, self.nq)), self._f(p1, (B, self.nd))
        u1 = None if (u1 is None or self.nu == 0) else self._f(u1, (B, self.nu))
        lam = None if (lambda1 is None or self.nc == 0) else self._f(lambda1, (B, self.nc))
        f = np.empty((B, self.nd + self.nc))
        self.calc_f_raw(False, B, float(t1), float(t2), q1, q2, p1, u1, lam, f)
        return f

    def discrete_fm2(self, t1, t2, q1, q2, u1=None):
        """Discrete forcing (t2 - t1) F(q_mid, dq, u1) [B][nd] (_MidpointVI.discrete_fm2, midpointvi.c:2710-2727)."""
        q1 = np.atleast_2d(np.asarray(q1, float))
        B = q1.shape[0]
        q1, q2 = self._f(q1, (B, self.nq)), self._f(q2, (B, self.nq))
        u1 = None if (u1 is None or self.nu == 0) else self._f(u1, (B, self.nu))
        fm2 = np.empty((B, self.nd))
        self.discrete_fm2_raw(False, B, float(t1), float(t2), q1, q2, u1, fm2)
        return fm2

    def step(self, q1, p1, t0, dt, nsteps=1, u1=None, k2=None, q2_guess=None, lambda_guess=None,
             tolerance=1e-10, max_iterations=200, sample_every=0, times=None):
        """`nsteps` consecutive MidpointVI steps for every instance.  Returns a dict with the
        final q2, p2, lambda1, iters (summed), status and optionally the sampled trajectory.
        times: optional grid [nsteps+1] shared by the batch (step s runs times[s] -> times[s+1])."""
        if times is not None:
            times = self._f(times, (nsteps + 1,))
        q1 = np.atleast_2d(np.asarray(q1, float))
        B = q1.shape[0]
        q1, p1 = self._f(q1, (B, self.nq)), self._f(p1, (B, self.nd))
        u1 = None if (u1 is None or self.nu == 0) else self._f(u1, (B, nsteps, self.nu))
        if self.nk:
            if k2 is None:
                raise ValueError("k2 [B][nsteps][nk] is required: the system has %d kinematic configs" % self.nk)
            k2 = self._f(k2, (B, nsteps, self.nk))
        else:
            k2 = None
        q2g = None if q2_guess is None else self._f(q2_guess, (B, self.nd))
        lg = None if (lambda_guess is None or self.nc == 0) else self._f(lambda_guess, (B, self.nc))
        out = dict(q2=np.empty((B, self.nq)), p2=np.empty((B, self.nd)),
                   lambda1=np.zeros((B, self.nc)), iters=np.zeros(B, np.int32),
                   status=np.zeros(B, np.int32))
        ns = nsteps // sample_every if sample_every > 0 else 0
        if ns:
            out["traj_q"] = np.empty((B, ns, self.nq))
            out["traj_p"] = np.empty((B, ns, self.nd))
        self.step_raw(False, B, nsteps, float(t0), float(dt), q1, p1, u1, k2, q2g, lg, out["q2"],
                      out["p2"], out["lambda1"] if self.nc else None, out["iters"], out["status"],
                      tolerance, max_iterations, sample_every, out.get("traj_q"), out.get("traj_p"), times=times)
        return out

    def project(self, bX, bU, Kfb, t0, dt, use_hint=True, tolerance=1e-10, max_iterations=200, times=None):
        """Closed-loop rollouts X[0] = bX[0], U[k] = bU[k] - K[k](X[k] - bX[k]), X[k+1] = f(X[k], U[k])
        (DSystem.project / DOptimizer.armijo_simulate) for a batch of candidates.
        bX [B,K+1,nX], bU [B,K,nU], Kfb [K,nU,nX] (shared) or [B,K,nU,nX]."""
        bX = np.ascontiguousarray(np.asarray(bX, float))
        if bX.ndim == 2:
            bX = bX[None]
        B, K = bX.shape[0], bX.shape[1] - 1
        bX = self._f(bX, (B, K + 1, self.nX))
        bU = self._f(bU, (B, K, self.nU))
        Kfb = np.ascontiguousarray(np.asarray(Kfb, float))
        per = Kfb.ndim == 4
        Kfb = self._f(Kfb, (B, K, self.nU, self.nX) if per else (K, self.nU, self.nX))
        out = dict(X=np.zeros((B, K + 1, self.nX)), U=np.zeros((B, K, self.nU)), iters=np.zeros(B, np.int32),
                   status=np.zeros(B, np.int32), fail_step=np.zeros(B, np.int32))
        self.project_raw(False, B, K, float(t0), float(dt), bX, bU, Kfb, out["X"], out["U"], out["status"],
                         k_per_instance=per, use_hint=use_hint, iters=out["iters"], fail_step=out["fail_step"],
                         tolerance=tolerance, max_iterations=max_iterations,
                         times=None if times is None else self._f(times, (K + 1,)))
        return out

    def linearize(self, q1, p1, u1=None, k2=None, t1=0.0, t2=None, dt=None, q2_guess=None,
                  lambda_guess=None, want_raw=False, tolerance=1e-10, max_iterations=200):
        """One DSystem-style linearization per instance: solve the step, then A (fdx) and B (fdu)."""
        q1 = np.atleast_2d(np.asarray(q1, float))
        B = q1.shape[0]
        q1, p1 = self._f(q1, (B, self.nq)), self._f(p1, (B, self.nd))
        u1 = self._f(u1 if u1 is not None else np.zeros((B, self.nu)), (B, self.nu))
        k2 = self._f(k2 if k2 is not None else np.zeros((B, self.nk)), (B, self.nk))
        q2g = None if q2_guess is None else self._f(q2_guess, (B, self.nd))
        lg = None if (lambda_guess is None or self.nc == 0) else self._f(lambda_guess, (B, self.nc))
        t1a = t2a = None
        t1s = dts = 0.0
        if np.ndim(t1) == 0 and (t2 is None or np.ndim(t2) == 0):
            t1s = float(t1)
            dts = float(dt) if t2 is None else float(t2) - float(t1)
            if t2 is not None:
                # keep the reference's dt = t2 - t1 rounding: pass both as arrays
                t1a = np.full(B, float(t1)); t2a = np.full(B, float(t2))
        else:
            t1a = self._f(np.broadcast_to(t1, (B,)), (B,))
            t2a = self._f(np.broadcast_to(t2, (B,)), (B,))
        out = dict(q2=np.empty((B, self.nq)), p2=np.empty((B, self.nd)), lambda1=np.zeros((B, self.nc)),
                   iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32),
                   A=np.empty((B, self.nX, self.nX)), B=np.zeros((B, self.nX, self.nU)))
        rawbufs = {}
        if want_raw:
            wrt = {"dq1": self.nq, "dp1": self.nd, "du1": self.nu, "dk2": self.nk}
            for n in RAW:
                rawbufs[n] = np.zeros((B, wrt[n[3:]], self.nc if n.startswith("l1") else self.nd))
            out.update(rawbufs)
        self.linearize_raw(False, B, q1, p1, u1 if self.nu else None, k2 if self.nk else None,
                           out["status"], t1=t1a, t2=t2a, t1_scalar=t1s, dt_scalar=dts, q2_guess=q2g,
                           lambda_guess=lg, q2=out["q2"], p2=out["p2"],
                           lambda1=out["lambda1"] if self.nc else None, iters=out["iters"],
                           A=out["A"], B=out["B"] if self.nU else None,
                           raw={k: v for k, v in rawbufs.items() if v.size},
                           tolerance=tolerance, max_iterations=max_iterations)
        return out

    def d2_shapes(self, B):
        cnt = {"dq1": self.nq, "dp1": self.nd, "du1": self.nu, "dk2": self.nk}
        out = {}
        for w in D2_WHICH:
            for kd in D2_KINDS:
                a, b = kd[:3], kd[3:]
                out[w + "_" + kd] = (B, cnt[a], cnt[b], self.nc if w == "l1" else self.nd)
        return out

    def deriv2(self, q1, p1, u1=None, k2=None, t1=0.0, t2=None, dt=None, q2_guess=None,
               lambda_guess=None, tolerance=1e-10, max_iterations=200, z=None, tensors=True):
        """solve + first derivatives + every second-derivative tensor (reference layout
        [B][wrt A][wrt B][out]); returns the dict of `linearize(want_raw=True)` plus the tensors."""
        q1 = np.atleast_2d(np.asarray(q1, float))
        B = q1.shape[0]
        q1, p1 = self._f(q1, (B, self.nq)), self._f(p1, (B, self.nd))
        u1 = self._f(u1 if u1 is not None else np.zeros((B, self.nu)), (B, self.nu))
        k2 = self._f(k2 if k2 is not None else np.zeros((B, self.nk)), (B, self.nk))
        q2g = None if q2_guess is None else self._f(q2_guess, (B, self.nd))
        lg = None if (lambda_guess is None or self.nc == 0) else self._f(lambda_guess, (B, self.nc))
        t1a = self._f(np.broadcast_to(t1, (B,)), (B,))
        t2a = self._f(np.broadcast_to(t1a + dt if t2 is None else t2, (B,)), (B,))
        out = dict(q2=np.empty((B, self.nq)), p2=np.empty((B, self.nd)), lambda1=np.zeros((B, self.nc)),
                   iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32),
                   A=np.empty((B, self.nX, self.nX)), B=np.zeros((B, self.nX, self.nU)))
        wrt = {"dq1": self.nq, "dp1": self.nd, "du1": self.nu, "dk2": self.nk}
        rawbufs = {n: np.zeros((B, wrt[n[3:]], self.nc if n.startswith("l1") else self.nd)) for n in RAW}
        d2 = {n: np.zeros(sh) for n, sh in self.d2_shapes(B).items()} if tensors else {}
        zk = {}
        if z is not None:
            zk = dict(z=self._f(z, (B, self.nX)), fdxdx=np.zeros((B, self.nX, self.nX)),
                      fdxdu=np.zeros((B, self.nX, self.nU)), fdudu=np.zeros((B, self.nU, self.nU)))
            out.update({k: v for k, v in zk.items() if k != "z"})
            if not self.nU:
                zk["fdxdu"] = zk["fdudu"] = None
        self.deriv2_raw(False, B, q1, p1, u1 if self.nu else None, k2 if self.nk else None, out["status"],
                        {n: v for n, v in d2.items() if v.size}, **zk, t1=t1a, t2=t2a, q2_guess=q2g,
                        lambda_guess=lg, q2=out["q2"], p2=out["p2"],
                        lambda1=out["lambda1"] if self.nc else None, iters=out["iters"], A=out["A"],
                        B=out["B"] if self.nU else None, raw={k: v for k, v in rawbufs.items() if v.size},
                        tolerance=tolerance, max_iterations=max_iterations)
        out.update(rawbufs)
        out.update(d2)
        return out
