"""Newton-iteration-count flip fraction against the reference's own C loop at a chosen scale (development aid; the
1e7-step version is tests/test_gpu_parity_r2.py::test_iteration_count_flip_fraction_at_1e7_steps).
    NAME=damped_pendulum B=102400 STEPS=1000 python tools/flip_fraction.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import golden_util as G
import cpu_baseline as cb
from trep_b200 import lib
name = os.environ.get("NAME", "damped_pendulum"); B = int(os.environ.get("B", "102400")); nsteps = int(os.environ.get("STEPS", "1000"))
rng = np.random.default_rng(2025)
d = G.desc(name)
q = rng.uniform(-np.pi, np.pi, (B, d.nq))
if name == "pend_on_cart1":
    q[:, 0] = rng.uniform(-1, 1, B)
p = rng.normal(0, 2.0, (B, d.nd))
want = cb.rollouts_parallel(name, q, p, nsteps, 0.0, 0.01)
got = lib.System(d).step(q, p, 0.0, 0.01, nsteps=nsteps)
assert np.array_equal(got["status"], want["status"])
ok = want["status"] == 0
diff = np.abs(got["iters"][ok].astype(np.int64) - want["iters"][ok])
err = np.maximum(np.max(np.abs(got["q2"][ok] - want["q2"][ok]), axis=1) / np.maximum(1.0, np.max(np.abs(want["q2"][ok]), axis=1)),
                 np.max(np.abs(got["p2"][ok] - want["p2"][ok]), axis=1) / np.maximum(1.0, np.max(np.abs(want["p2"][ok]), axis=1)))
print("%s: %d rollouts x %d steps = %.2e DEL steps, ok %.5f; rollouts with a different iteration total %d, flip fraction %.2e per step; "
      "final-state error median %.2e, 99.9th percentile %.2e, max %.2e" % (name, B, nsteps, ok.sum() * nsteps, ok.mean(), int((diff > 0).sum()),
      diff.sum() / float(ok.sum() * nsteps), np.median(err), np.quantile(err, 0.999), err.max()))
