"""Phase breakdown of the general linearize kernel on the marionette (needs libtrepb_prof.so:
python -m trep_b200.build --prof; run with TREPB_LIBPATH=trep_b200/libtrepb_prof.so)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(a)
rng = np.random.default_rng(0)
B = int(os.environ.get("PUPPET_B", "32768"))
d = systems.named_desc("puppet"); s = lib.System(d, specialize=os.environ.get("SPEC", "1") == "1", cooperative={"1": True, "0": False}.get(os.environ.get("COOP", ""), None))
g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
idx = rng.integers(1, 58, B)
q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
dq, dp, dk, dl = up(q1), up(p1), up(g["roll_k2"][idx]), up(g["roll_lambda"][idx - 1])
q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd)); l2 = lib.DeviceBuffer(0, (B, d.nc))
it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
A = lib.DeviceBuffer(0, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (B, d.nX, d.nU))
ticks = (C.c_ulonglong * 32)()
names = {0: "solve: Dh(q1)", 1: "solve: residual eval_mid(1)", 2: "solve: h(q2)", 3: "solve: eval_mid_again(2)",
         4: "solve: Jacobian assembly + Dh(q2)", 5: "solve: LU 28x28 + solve", 8: "deriv1: constraints q1 (DDh.lam), q2",
         9: "deriv1: eval_mid_again(2)", 10: "deriv1: table assembly", 11: "deriv1: M2 LU, proj", 12: "deriv1: 80 rhs solves + A/B writes",
         13: "deriv1: constant blocks of A,B",
         16: "coop solve: pose(q1) + Dh1", 17: "coop solve: pose(q2) + h + Dh2", 18: "coop solve: pose(mid) + V",
         19: "coop solve: dyn_first (inertia, up sweep, Lq/Lv)", 20: "coop solve: residual + test",
         21: "coop solve: dyn_second (H,G,P + pair tables)", 22: "coop solve: Jacobian assembly",
         23: "coop solve: LU 28x28(+1)", 24: "coop solve: back substitution + update",
         25: "coop deriv1: dyn_second", 26: "coop deriv1: pose(q1) + DDh.lambda", 27: "coop deriv1: M2 / T22 assembly",
         28: "coop deriv1: LU M2(+Dh1^T), proj, aux", 29: "coop deriv1: rhs columns + outputs", 30: "coop deriv1: constant blocks of A,B"}
for rep in range(2):
    lib.raw().trepb_phase_ticks(ticks, 1)
    s.linearize_raw(True, B, dq, dp, None, dk, st, t1_scalar=0.0, dt_scalar=0.01, lambda_guess=dl, q2=q2, p2=p2,
                    lambda1=l2, iters=it, A=A, B=Bm)
    lib.synchronize(0)
lib.raw().trepb_phase_ticks(ticks, 0)
tot = sum(ticks)
print("B=%d kernel %.2f ms  %.3e lin/s" % (B, s.last_kernel_ms(), B / s.last_kernel_ms() * 1e3))
for i in range(32):
    if ticks[i]:
        print("  %-45s %5.1f %%" % (names.get(i, str(i)), 100.0 * ticks[i] / tot))
