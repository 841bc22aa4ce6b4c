import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(np.ascontiguousarray(a))
rng = np.random.default_rng(0)
d = systems.named_desc("pccd"); B = 1 << 14
g = np.load("tests/golden/pccd.npz")
idx = rng.integers(1, g["roll_q"].shape[0] - 1, B)
q1 = g["roll_q"][idx] + rng.normal(0, 0.01, (B, d.nq)); p1 = g["roll_p"][idx] + rng.normal(0, 0.05, (B, d.nd))
for label, kw in (("default", {}), ("rt", dict(specialize=False))):
    s = lib.System(d, **kw)
    dq, dp, lam = up(q1), up(p1), up(g["roll_lambda"][idx - 1])
    st = lib.DeviceBuffer(0, (B,), np.int32)
    z = up(rng.normal(0, 1, (B, d.nX))); xx = lib.DeviceBuffer(0, (B, d.nX, d.nX))
    A = lib.DeviceBuffer(0, (B, d.nX, d.nX))
    for rep in range(3):
        s.linearize_raw(True, B, dq, dp, None, None, st, t1_scalar=0.0, dt_scalar=0.01, A=A, lambda_guess=lam)
        lib.synchronize(0); ml = s.last_kernel_ms()
        lib.synchronize(0); t0 = time.perf_counter()
        s.deriv2_raw(True, B, dq, dp, None, None, st, {}, z=z, fdxdx=xx, t1_scalar=0.0, dt_scalar=0.01, lambda_guess=lam)
        lib.synchronize(0); tot = (time.perf_counter() - t0) * 1e3; md = s.last_kernel_ms()
    print(label, s.kernel_name, "lin %.2f ms, d2 passes %.2f ms, deriv2 call total %.2f ms" % (ml, md, tot))
    s.close()
