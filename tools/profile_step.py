"""One short launch of each hot kernel, for ncu (development/profiling aid).
usage: python tools/profile_step.py [step|lin|puppet|d2|all]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(a)
rng = np.random.default_rng(0)

if mode in ("step", "all"):
    B, nsteps = 1 << 20, 50
    d = systems.named_desc("damped_pendulum"); s = lib.System(d)
    th0 = rng.uniform(-np.pi, np.pi, (B, 1)); th1 = th0 + rng.uniform(-0.02, 0.02, (B, 1))
    dq0, dq1 = up(th0), up(th1); dp = lib.DeviceBuffer(0, (B, 1))
    s.calc_p2_raw(True, B, 0.01, dq0, dq1, dp)
    q2 = lib.DeviceBuffer(0, (B, 1)); p2 = lib.DeviceBuffer(0, (B, 1))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    for _ in range(2):
        s.step_raw(True, B, nsteps, 0.01, 0.01, dq1, dp, None, None, None, None, q2, p2, None, it, st)
    lib.synchronize(0)
    print("step: B=%d nsteps=%d ms=%.3f iters/step=%.3f" % (B, nsteps, s.last_kernel_ms(), it.download().mean() / nsteps))

if mode in ("pend5", "all"):
    B, nsteps = 1 << 18, 20
    d = systems.named_desc("pendulum5"); s = lib.System(d)
    q0 = np.zeros((B, 5)); q0[:, 0] = rng.uniform(-np.pi, np.pi, B)
    dq = up(q0); dp = lib.DeviceBuffer(0, (B, 5))
    s.calc_p2_raw(True, B, 0.01, dq, dq, dp)
    q2 = lib.DeviceBuffer(0, (B, 5)); p2 = lib.DeviceBuffer(0, (B, 5))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    for _ in range(2):
        s.step_raw(True, B, nsteps, 0.01, 0.01, dq, dp, None, None, None, None, q2, p2, None, it, st)
    lib.synchronize(0)
    print("pend5: B=%d nsteps=%d ms=%.3f iters/step=%.3f" % (B, nsteps, s.last_kernel_ms(), it.download().mean() / nsteps))

if mode in ("dual", "all"):
    B, nsteps = 1 << 20, 50
    d = systems.named_desc("dual_pendulums"); s = lib.System(d)
    dq = up(rng.uniform(-np.pi, np.pi, (B, 2))); dp = lib.DeviceBuffer(0, (B, 2))
    s.calc_p2_raw(True, B, 0.01, dq, dq, dp)
    q2 = lib.DeviceBuffer(0, (B, 2)); p2 = lib.DeviceBuffer(0, (B, 2))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    for _ in range(2):
        s.step_raw(True, B, nsteps, 0.01, 0.01, dq, dp, None, None, None, None, q2, p2, None, it, st)
    lib.synchronize(0)
    print("dual: B=%d nsteps=%d ms=%.3f iters/step=%.3f" % (B, nsteps, s.last_kernel_ms(), it.download().mean() / nsteps))

if mode in ("lin", "all"):
    B = 1 << 22
    d = systems.named_desc("pend_on_cart1"); s = lib.System(d)
    dq, dp, du = up(rng.uniform(-0.5, 0.5, (B, 2))), up(rng.normal(0, 1, (B, 2))), up(rng.uniform(-2, 2, (B, 1)))
    q2 = lib.DeviceBuffer(0, (B, 2)); p2 = lib.DeviceBuffer(0, (B, 2))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    A = lib.DeviceBuffer(0, (B, 4, 4)); Bm = lib.DeviceBuffer(0, (B, 4, 1))
    for _ in range(2):
        s.linearize_raw(True, B, dq, dp, du, None, st, t1_scalar=0.0, dt_scalar=0.01, q2=q2, p2=p2, iters=it, A=A, B=Bm)
    lib.synchronize(0)
    print("lin: B=%d ms=%.3f" % (B, s.last_kernel_ms()))

if mode in ("puppet", "all"):
    B = int(os.environ.get("PUPPET_B", "8192"))
    d = systems.named_desc("puppet"); s = lib.System(d)
    g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
    idx = rng.integers(1, 58, B)
    q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
    q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
    dq, dp, dk, dl = up(q1), up(p1), up(g["roll_k2"][idx]), up(g["roll_lambda"][idx - 1])
    q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd)); l2 = lib.DeviceBuffer(0, (B, d.nc))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    A = lib.DeviceBuffer(0, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (B, d.nX, d.nU))
    for _ in range(2):
        s.linearize_raw(True, B, dq, dp, None, dk, st, t1_scalar=0.0, dt_scalar=0.01, lambda_guess=dl, q2=q2, p2=p2,
                        lambda1=l2, iters=it, A=A, B=Bm)
    lib.synchronize(0)
    print("puppet lin: B=%d ms=%.3f iters=%.2f" % (B, s.last_kernel_ms(), it.download().mean()))

if mode in ("d2", "all"):
    B = int(os.environ.get("D2_B", "128"))
    d = systems.named_desc("puppet"); s = lib.System(d)
    g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
    idx = rng.integers(1, 58, B)
    q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
    q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
    dq, dp, dk, dl = up(q1), up(p1), up(g["roll_k2"][idx]), up(g["roll_lambda"][idx - 1])
    st = lib.DeviceBuffer(0, (B,), np.int32)
    z = up(rng.normal(0, 1, (B, d.nX)))
    xx = lib.DeviceBuffer(0, (B, d.nX, d.nX)); xu = lib.DeviceBuffer(0, (B, d.nX, d.nU)); uu = lib.DeviceBuffer(0, (B, d.nU, d.nU))
    for _ in range(2):
        s.deriv2_raw(True, B, dq, dp, None, dk, st, {}, z=z, fdxdx=xx, fdxdu=xu, fdudu=uu, t1_scalar=0.0, dt_scalar=0.01, lambda_guess=dl)
    lib.synchronize(0)
    print("puppet d2: B=%d ms=%.3f" % (B, s.last_kernel_ms()))
