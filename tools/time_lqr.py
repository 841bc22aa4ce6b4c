"""Riccati sweep timing (development aid)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(a)
rng = np.random.default_rng(0)
nX, nU = int(os.environ.get("NX", "80")), int(os.environ.get("NU", "18"))
R, K = int(os.environ.get("R", "148")), int(os.environ.get("K", "32"))
A = up(np.eye(nX)[None, None] + rng.normal(0, 0.03, (R, K, nX, nX))); B = up(rng.normal(0, 1.0, (R, K, nX, nU)))
Q, Rm = up(np.eye(nX)), up(np.eye(nU))
Ko = lib.DeviceBuffer(0, (R, K, nU, nX)); st = lib.DeviceBuffer(0, (R,), np.int32)
for rep in range(3):
    lib.synchronize(0); t0 = time.perf_counter()
    lib.lqr_raw(True, 0, R, K, nX, nU, A, B, Q, Rm, Ko, st)
    lib.synchronize(0); dt = time.perf_counter() - t0
    ms = lib.lqr_last_kernel_ms(0)
    print("R=%d K=%d nX=%d nU=%d: host %.2f ms, kernel (CUDA events) %.2f ms, %.1f us per step per rollout-SM, ok=%s" % (R, K, nX, nU, dt * 1e3, ms, ms * 1e3 / K / max(1, (R + 147) // 148), bool(np.all(st.download() == 0))))
