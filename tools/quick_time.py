"""Quick device-side timing of the kernels (development aid; bench.py is the contract)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trep_b200 import lib, systems

def dbuf(a):
    return lib.DeviceBuffer(0, a.shape, a.dtype).upload(a)

print("fp64 peak TFLOP/s:", lib.measure_fp64_peak(0))
rng = np.random.default_rng(0)
for name, B, nsteps in (("damped_pendulum", 1 << 20, 100), ("pendulum1", 1 << 20, 100), ("dual_pendulums", 1 << 20, 50),
                        ("pend_on_cart1", 1 << 20, 50), ("pendulum5", 1 << 18, 20)):
    for spec in (True, False):
        d = systems.named_desc(name)
        s = lib.System(d, specialize=spec)
        q1 = rng.uniform(-3, 3, (B, d.nq)); p1 = rng.normal(0, 1, (B, d.nd))
        dq, dp = dbuf(q1), dbuf(p1)
        du = dbuf(np.zeros((B, nsteps, d.nu))) if d.nu else None
        q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd))
        it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
        for rep in range(2):
            s.step_raw(True, B, nsteps, 0.0, 0.01, dq, dp, du, None, None, None, q2, p2, None, it, st)
            lib.synchronize(0)
        ms = s.last_kernel_ms()
        iters = it.download()
        print("%-16s %-8s step: B=%d nsteps=%d  %.2f ms  %.3e steps/s  mean iters/step %.2f  info %s" % (
            name, s.kernel_name[:8], B, nsteps, ms, B * nsteps / ms * 1e3, iters.mean() / nsteps, s.kernel_info(0)))
        # linearize
        Bl = B
        du1 = dbuf(rng.uniform(-1, 1, (Bl, d.nu))) if d.nu else None
        A = lib.DeviceBuffer(0, (Bl, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (Bl, d.nX, max(d.nU, 1)))
        for rep in range(2):
            s.linearize_raw(True, Bl, dq, dp, du1, None, st, t1_scalar=0.0, dt_scalar=0.01, q2=q2, p2=p2, iters=it, A=A, B=Bm if d.nU else None)
            lib.synchronize(0)
        ms = s.last_kernel_ms()
        byt = Bl * 8 * (d.nq + d.nd + d.nu + d.nX * d.nX + d.nX * d.nU + d.nq + d.nd + 1)
        print("%-16s %-8s lin : B=%d  %.2f ms  %.3e lin/s  %.1f GB/s  info %s" % (
            name, s.kernel_name[:8], Bl, ms, Bl / ms * 1e3, byt / ms / 1e6, s.kernel_info(2)))
        for b in (dq, dp, q2, p2, it, st, A, Bm): b.free()
        s.close()

# puppet
d = systems.named_desc("puppet")
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "puppet.npz"))
s = lib.System(d)
for B in (4096, 32768):
    idx = rng.integers(1, 58, B)
    q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
    q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
    k2 = g["roll_k2"][idx]; lam = g["roll_lambda"][idx - 1]
    dq, dp, dk, dl = dbuf(q1), dbuf(p1), dbuf(k2), dbuf(lam)
    q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd)); l2 = lib.DeviceBuffer(0, (B, d.nc))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    A = lib.DeviceBuffer(0, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (B, d.nX, d.nU))
    for rep in range(2):
        s.step_raw(True, B, 1, 0.0, 0.01, dq, dp, None, dk, None, dl, q2, p2, l2, it, st)
        lib.synchronize(0)
    ms = s.last_kernel_ms()
    print("puppet step: B=%d %.2f ms %.3e steps/s iters %.2f status ok %s info %s" % (B, ms, B / ms * 1e3, it.download().mean(), (st.download() == 0).mean(), s.kernel_info(0)))
    for rep in range(2):
        s.linearize_raw(True, B, dq, dp, None, dk, st, t1_scalar=0.0, dt_scalar=0.01, lambda_guess=dl, q2=q2, p2=p2, lambda1=l2, iters=it, A=A, B=Bm)
        lib.synchronize(0)
    ms = s.last_kernel_ms()
    print("puppet lin : B=%d %.2f ms %.3e lin/s status ok %s info %s" % (B, ms, B / ms * 1e3, (st.download() == 0).mean(), s.kernel_info(2)))
    for b in (dq, dp, dk, dl, q2, p2, l2, it, st, A, Bm): b.free()
