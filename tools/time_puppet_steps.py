"""Marionette in-kernel DEL steps/s (development aid; TREPB_COOP_WARPS caps the instances per SM)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(np.ascontiguousarray(a))
rng = np.random.default_rng(0)
d = systems.named_desc("puppet"); s = lib.System(d)
g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
B, ns = int(os.environ.get("B", "32768")), 16
idx = rng.integers(1, 40, B)
q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
k2 = np.repeat(g["roll_k2"][idx][:, None, :], ns, axis=1)
dq, dp, dk, dl = up(q1), up(p1), up(k2), up(g["roll_lambda"][idx - 1])
q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd)); l2 = lib.DeviceBuffer(0, (B, d.nc))
it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
for rep in range(3):
    s.step_raw(True, B, ns, 0.0, 0.01, dq, dp, None, dk, None, dl, q2, p2, l2, it, st)
    lib.synchronize(0)
ms = s.last_kernel_ms()
print("%s B=%d x %d steps: %.2f ms  %.3e DEL steps/s  iters/step %.2f ok=%.3f info=%s" % (
    s.kernel_name, B, ns, ms, B * ns / ms * 1e3, it.download().mean() / ns, (st.download() == 0).mean(), s.kernel_info(0)))
