"""Device-side timing of the second-derivative path (development aid)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(a)
rng = np.random.default_rng(0)
CASES = (("damped_pendulum", 1 << 20), ("pend_on_cart1", 1 << 20), ("puppet", int(os.environ.get("PUPPET_B", "256"))))
if os.environ.get("ONLY"):
    CASES = tuple(c for c in CASES if c[0] == os.environ["ONLY"])
for name, B in CASES:
    d = systems.named_desc(name); s = lib.System(d, d2_pairwise=bool(int(os.environ.get("PAIRWISE", "0"))))
    if name == "puppet":
        g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
        idx = rng.integers(1, 58, B)
        q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
        q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
        k2 = up(g["roll_k2"][idx]); lam = up(g["roll_lambda"][idx - 1]); u1 = None
    else:
        q1 = rng.uniform(-3, 3, (B, d.nq)); p1 = rng.normal(0, 1, (B, d.nd))
        u1 = up(rng.uniform(-1, 1, (B, d.nu))) if d.nu else None; k2 = None; lam = None
    dq, dp = up(q1), up(p1)
    st = lib.DeviceBuffer(0, (B,), np.int32)
    d2 = {n: lib.DeviceBuffer(0, sh) for n, sh in s.d2_shapes(B).items() if int(np.prod(sh))}
    t0 = time.perf_counter()
    for rep in range(2):
        s.deriv2_raw(True, B, dq, dp, u1, k2, st, d2, t1_scalar=0.0, dt_scalar=0.01, lambda_guess=lam)
        lib.synchronize(0)
        ms = s.last_kernel_ms()
    nx = d.nq + d.nd + d.nu + d.nk
    print("%-16s B=%d pairs=%d d2 kernel %.2f ms -> %.3e deriv2/s (kernel only), ok=%.3f" % (
        name, B, nx * (nx + 1) // 2, ms, B / ms * 1e3, (st.download() == 0).mean()))
