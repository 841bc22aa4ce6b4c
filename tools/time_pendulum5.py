import sys; sys.path.insert(0, ".")
import numpy as np
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(np.ascontiguousarray(a))
rng = np.random.default_rng(0)
d = systems.named_desc("pendulum5")
B, nsteps = 1 << 18, 200
q0 = np.zeros((B, 5)); q0[:, 0] = rng.uniform(-np.pi, np.pi, B)
for label, kw in (("spec", {}), ("general", dict(specialize=False, cooperative=False)), ("coop", dict(specialize=False, cooperative=True)), ("coop-default", dict(cooperative=True))):
    s = lib.System(d, **kw)
    dq = up(q0); dp = lib.DeviceBuffer(0, (B, 5))
    s.calc_p2_raw(True, B, 0.01, dq, dq, dp)
    q2 = lib.DeviceBuffer(0, (B, 5)); p2 = lib.DeviceBuffer(0, (B, 5))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    for rep in range(2):
        s.step_raw(True, B, nsteps, 0.01, 0.01, dq, dp, None, None, None, None, q2, p2, None, it, st)
        lib.synchronize(0)
        ms = s.last_kernel_ms()
    print("%-12s kernel=%-20s %.2f ms -> %.3e steps/s iters/step %.3f" % (label, s.kernel_name, ms, B * nsteps / ms * 1e3, it.download().mean() / nsteps))
    for b in (dq, dp, q2, p2, it, st): b.free()
    s.close()
