"""Marionette linearize/step timing at several batch sizes (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(a)
rng = np.random.default_rng(0)
d = systems.named_desc("puppet"); s = lib.System(d, specialize=os.environ.get("SPEC", "1") == "1", cooperative={"1": True, "0": False}.get(os.environ.get("COOP", ""), None))
print("kernel:", s.kernel_name)
g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
for B in [int(x) for x in os.environ.get("BS", "32768,131072").split(",")]:
    idx = rng.integers(1, 58, B)
    q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
    q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
    dq, dp, dk, dl = up(q1), up(p1), up(g["roll_k2"][idx]), up(g["roll_lambda"][idx - 1])
    q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd)); l2 = lib.DeviceBuffer(0, (B, d.nc))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    A = lib.DeviceBuffer(0, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (B, d.nX, d.nU))
    for rep in range(3):
        s.linearize_raw(True, B, dq, dp, None, dk, st, t1_scalar=0.0, dt_scalar=0.01, lambda_guess=dl, q2=q2, p2=p2,
                        lambda1=l2, iters=it, A=A, B=Bm)
        lib.synchronize(0)
    ms = s.last_kernel_ms()
    print("%s B=%d lin %.2f ms %.3e lin/s ok=%.3f info=%s" % (os.environ.get("TREPB_LIBPATH", "default")[-14:], B, ms, B / ms * 1e3, (st.download() == 0).mean(), s.kernel_info(2)))
    for b in (dq, dp, dk, dl, q2, p2, l2, it, st, A, Bm): b.free()
