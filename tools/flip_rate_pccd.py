import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import numpy as np
import ref_systems as R
import golden_util as G
from trep_b200 import lib
rng = np.random.default_rng(21)
system, mvi = R.make_mvi("pccd")
nq, nd = mvi.nq, mvi.nd
B = 3000
g = G.golden("pccd")
idx = rng.integers(1, g["roll_q"].shape[0] - 1, B)
q1 = g["roll_q"][idx] + rng.normal(0, 0.01, (B, nq)); p1 = g["roll_p"][idx] + rng.normal(0, 0.05, (B, nd))
lam = g["roll_lambda"][idx - 1]
u1 = np.zeros((B, 0)); k2 = np.zeros((B, 0)); t1 = np.zeros(B); t2 = t1 + 0.01
want = R.run_cases(mvi, t1, t2, q1, p1, u1, k2, lambda_guess=lam, deriv1=False)
for label, kw in (("general", dict(specialize=False, cooperative=False)), ("coop", dict(specialize=False, cooperative=True)), ("coop-static", {})):
    s = lib.System(G.desc("pccd"), **kw)
    out = s.linearize(q1, p1, u1, k2, t1=t1, t2=t2, lambda_guess=lam)
    ok = (out["status"] == 0) & (want["status"] == 0)
    print(label, s.kernel_name, "status equal:", np.array_equal(out["status"], want["status"]), "flips:", int(np.sum(out["iters"][ok] != want["iters"][ok])), "of", int(ok.sum()),
          "max rel err q2 %.2e" % (np.max(np.abs(out["q2"][ok] - want["q2"][ok])) / max(1, np.max(np.abs(want["q2"][ok])))), "mean iters", out["iters"][ok].mean(), want["iters"][ok].mean())
