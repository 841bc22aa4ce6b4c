"""Linearizations/s of the table-driven thread kernel vs the cooperative kernel on mid-size systems
(development aid: where should the library switch from one thread to one warp per instance?)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(np.ascontiguousarray(a))
rng = np.random.default_rng(0)
for name, B in (("pendulum5", 1 << 18), ("pccd", 1 << 16), ("pend_on_cart2", 1 << 20)):
    d = systems.named_desc(name)
    for label, kw in (("general", dict(specialize=False, cooperative=False)), ("coop", dict(specialize=False, cooperative=True)),
                      ("coop-ct", dict(specialize=True, cooperative=True))):
        try:
            s = lib.System(d, **kw)
        except lib.TrepbError as e:
            print("%-14s %-8s not available: %s" % (name, label, str(e)[:80]))
            continue
        lam = None
        if name == "pccd":
            g = np.load(os.path.join(ROOT, "tests", "golden", "pccd.npz"))
            idx = rng.integers(1, g["roll_q"].shape[0] - 1, B)
            q1 = g["roll_q"][idx] + rng.normal(0, 0.01, (B, d.nq)); p1 = g["roll_p"][idx] + rng.normal(0, 0.05, (B, d.nd))
            lam = up(g["roll_lambda"][idx - 1])
        else:
            q1 = rng.uniform(-3, 3, (B, d.nq)); p1 = rng.normal(0, 1, (B, d.nd))
        dq, dp = up(q1), up(p1)
        du = up(rng.uniform(-1, 1, (B, d.nu))) if d.nu else None
        st = lib.DeviceBuffer(0, (B,), np.int32)
        A = lib.DeviceBuffer(0, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (B, d.nX, d.nU)) if d.nU else None
        for rep in range(3):
            s.linearize_raw(True, B, dq, dp, du, None, st, t1_scalar=0.0, dt_scalar=0.01, A=A, B=Bm, lambda_guess=lam)
            lib.synchronize(0)
            ms = s.last_kernel_ms()
        print("%-14s %-8s kernel=%-18s B=%d  %.3f ms -> %.3e lin/s  ok=%.4f  %s" % (name, label, s.kernel_name, B, ms, B / ms * 1e3, (st.download() == 0).mean(), s.kernel_info(2)))
        for b in (dq, dp, du, st, A, Bm, lam):
            if b is not None: b.free()
        s.close()
