"""SM clock / power while the marionette linearize kernel runs back to back (development aid)."""
import os, subprocess, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(a)
rng = np.random.default_rng(0)
d = systems.named_desc("puppet"); s = lib.System(d)
g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
B = 131072
idx = rng.integers(1, 58, B)
q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
dq, dp, dk, dl = up(q1), up(p1), up(g["roll_k2"][idx]), up(g["roll_lambda"][idx - 1])
q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd)); l2 = lib.DeviceBuffer(0, (B, d.nc))
it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
A = lib.DeviceBuffer(0, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (B, d.nX, d.nU))
lines = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu",
                         "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l.strip()) for l in proc.stdout], daemon=True).start()
ms = []
for rep in range(40):
    s.linearize_raw(True, B, dq, dp, None, dk, st, t1_scalar=0.0, dt_scalar=0.01, lambda_guess=dl, q2=q2, p2=p2,
                    lambda1=l2, iters=it, A=A, B=Bm)
    lib.synchronize(0)
    ms.append(s.last_kernel_ms())
proc.terminate()
print("kernel ms first/last:", ms[0], ms[-1], "lin/s last: %.3e" % (B / ms[-1] * 1e3))
print("\n".join(lines[::3]))
