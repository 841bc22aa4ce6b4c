"""Linearizations/s and DEL steps/s of the table-driven thread kernel vs the cooperative kernels on small and
mid-size systems (development aid: where should the library switch from one thread to one warp per instance?)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from trep_b200 import lib, systems
import golden_util as G
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(np.ascontiguousarray(a))
rng = np.random.default_rng(0)
NAMES = os.environ.get("NAMES", "pend_on_cart2,wrench_arm,fourbar,rod,pendulum5,loop3d,spring_arms,pccd").split(",")
for name in NAMES:
    d = G.desc(name)
    B = int(os.environ.get("B", str(1 << 17)))
    for label, kw in (("general", dict(specialize=False, cooperative=False)), ("coop", dict(specialize=False, cooperative=True)),
                      ("coop-ct", dict(specialize=True, cooperative=True)), ("default", dict(specialize=False))):
        try:
            s = lib.System(d, **kw)
        except lib.TrepbError as e:
            print("%-14s %-8s not available: %s" % (name, label, str(e)[:80]))
            continue
        if label == "coop-ct" and s.kernel_name == "cooperative":
            s.close(); continue
        if label == "default":
            print("%-14s default  kernel=%s" % (name, s.kernel_name)); s.close(); continue
        lam = None
        g = G.golden(name)
        if "roll_q" in g.files and g["roll_q"].shape[0] > 20:
            idx = rng.integers(1, g["roll_q"].shape[0] - 1, B)
            q1 = g["roll_q"][idx] + rng.normal(0, 0.01, (B, d.nq)); p1 = g["roll_p"][idx] + rng.normal(0, 0.05, (B, d.nd))
            if d.nc: lam = up(g["roll_lambda"][idx - 1])
            k2 = g["roll_k2"][idx] if d.nk else None
        else:
            q1 = rng.uniform(-3, 3, (B, d.nq)); p1 = rng.normal(0, 1, (B, d.nd)); k2 = None
        dq, dp = up(q1), up(p1)
        du = up(rng.uniform(-1, 1, (B, d.nu))) if d.nu else None
        dk = up(k2) if k2 is not None else None
        st = lib.DeviceBuffer(0, (B,), np.int32)
        A = lib.DeviceBuffer(0, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (B, d.nX, d.nU)) if d.nU else None
        for rep in range(3):
            s.linearize_raw(True, B, dq, dp, du, dk, st, t1_scalar=0.0, dt_scalar=0.01, A=A, B=Bm, lambda_guess=lam)
            lib.synchronize(0)
            ms = s.last_kernel_ms()
        ok = (st.download() == 0).mean()
        # 8 in-kernel steps with the inputs held
        ns = 8
        duu = up(np.zeros((B, ns, d.nu))) if d.nu else None
        dkk = up(np.repeat(k2[:, None, :], ns, axis=1)) if k2 is not None else None
        q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd)); it = lib.DeviceBuffer(0, (B,), np.int32)
        l2 = lib.DeviceBuffer(0, (B, d.nc)) if d.nc else None
        for rep in range(2):
            s.step_raw(True, B, ns, 0.0, 0.01, dq, dp, duu, dkk, None, lam, q2, p2, l2, it, st)
            lib.synchronize(0)
            ms2 = s.last_kernel_ms()
        print("%-14s %-8s kernel=%-18s B=%d  lin %.3e /s (ok %.3f)   steps %.3e /s (ok %.3f)  lin %s" % (
            name, label, s.kernel_name, B, B / ms * 1e3, ok, B * ns / ms2 * 1e3, (st.download() == 0).mean(),
            {k: v for k, v in s.kernel_info(2).items() if k in ("regs", "block")}))
        for b in (dq, dp, du, dk, st, A, Bm, lam, duu, dkk, q2, p2, it, l2):
            if b is not None: b.free()
        s.close()
