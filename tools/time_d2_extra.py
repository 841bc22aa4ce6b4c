"""Second-derivative throughput on the systems beyond BASELINE.json's configs (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(np.ascontiguousarray(a))
rng = np.random.default_rng(0)
for name, B in (("pccd", 1 << 14), ("wrench_arm", 1 << 17), ("pend_on_cart2", 1 << 18)):
    d = systems.named_desc(name)
    for label, kw in (("default", {}), ("general/jac", dict(specialize=False)), ("general/pairwise", dict(specialize=False, d2_pairwise=True))):
        s = lib.System(d, **kw)
        lam = None
        if name == "pccd":
            g = np.load(os.path.join(ROOT, "tests", "golden", "pccd.npz"))
            idx = rng.integers(1, g["roll_q"].shape[0] - 1, B)
            q1 = g["roll_q"][idx] + rng.normal(0, 0.01, (B, d.nq)); p1 = g["roll_p"][idx] + rng.normal(0, 0.05, (B, d.nd))
            lam = up(g["roll_lambda"][idx - 1])
        else:
            q1 = rng.uniform(-1, 1, (B, d.nq)); p1 = rng.normal(0, 1, (B, d.nd))
        dq, dp = up(q1), up(p1)
        du = up(rng.uniform(-1, 1, (B, d.nu))) if d.nu else None
        st = lib.DeviceBuffer(0, (B,), np.int32)
        z = up(rng.normal(0, 1, (B, d.nX)))
        xx = lib.DeviceBuffer(0, (B, d.nX, d.nX)); xu = lib.DeviceBuffer(0, (B, d.nX, max(d.nU, 1))); uu = lib.DeviceBuffer(0, (B, max(d.nU, 1), max(d.nU, 1)))
        import time
        for rep in range(3):
            lib.synchronize(0); t0 = time.perf_counter()
            s.deriv2_raw(True, B, dq, dp, du, None, st, {}, z=z, fdxdx=xx, fdxdu=xu if d.nU else None, fdudu=uu if d.nU else None,
                         t1_scalar=0.0, dt_scalar=0.01, lambda_guess=lam)
            lib.synchronize(0); ms = (time.perf_counter() - t0) * 1e3
        nx = d.nq + d.nd + d.nu + d.nk
        print("%-14s %-17s kernel=%-16s B=%d pairs=%d  %.2f ms (linearize + d2, host-timed) -> %.3e evaluations/s ok=%.3f" % (
            name, label, s.kernel_name, B, nx * (nx + 1) // 2, ms, B / ms * 1e3, (st.download() == 0).mean()))
        for b in (dq, dp, du, st, z, xx, xu, uu, lam):
            if b is not None: b.free()
        s.close()
