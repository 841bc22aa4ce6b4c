"""Operation-counted ("algorithmic") flops of the hot path's units, written to profiles/opcounts.json:
trepb_math.cuh - the math the thread-per-instance kernels run - compiled on the host with `double` replaced by a
counting wrapper (tests/opcount.cc), on samples of bench.py's workloads.
    python tools/count_ops.py
flops = add + sub + mul + div + sqrt + 2 fma; sin / cos evaluations are listed separately.  A division by dt is
executed as a multiplication by 1/dt and two fma corrections (div_dt, trepb_math.cuh), counted as executed: 5."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostmath as H
import opcount as OC
from trep_b200 import systems

DT = 0.01
out = {"definition": __doc__.split("\n\n")[0].replace("\n", " ") + " flops = add + sub + mul + div + sqrt + 2 fma; sin and cos "
       "evaluations separate; comparisons not counted."}


def acc(total, c):
    for k, v in c.items():
        total[k] = total.get(k, 0) + v


# ---- W2: damped pendulum, theta0 ~ U(-pi, pi), theta1 = theta0 + U(-0.02, 0.02), 1000 steps (bench.py workload_inputs)
rng = np.random.default_rng(0)
d = systems.named_desc("damped_pendulum")
n, steps = 64, 1000
th0 = rng.uniform(-np.pi, np.pi, (n, 1)); th1 = th0 + rng.uniform(-0.02, 0.02, (n, 1))
tot, its = {}, 0
for i in range(n):
    p1 = H.calc_p2(d, DT, th0[i], th1[i])
    r = OC.step(d, steps, 0.0, DT, th1[i], p1)
    assert r["iters"] >= 0
    acc(tot, r["counts"]); its += r["iters"]
chk = OC.step(d, 50, 0.0, DT, [0.7], [0.3])
out["damped_pendulum_step"] = {"unit": "DEL step", "sample": "%d rollouts x %d steps of W2" % (n, steps),
                               "flops": OC.flops(tot) / (n * steps), "sincos": tot["sincos"] / (n * steps),
                               "newton_iters": its / (n * steps), "counts_per_unit": {k: v / (n * steps) for k, v in tot.items()},
                               "check": {"iters": chk["iters"], "flops": OC.flops(chk["counts"]), "sincos": chk["counts"]["sincos"]}}

# ---- W4: dual pendulums, (theta1, theta2) ~ U(-pi, pi)^2 from rest, 100 steps
d = systems.named_desc("dual_pendulums")
n, steps = 64, 100
q = rng.uniform(-np.pi, np.pi, (n, 2))
tot, its = {}, 0
for i in range(n):
    p1 = H.calc_p2(d, DT, q[i], q[i])
    r = OC.step(d, steps, 0.0, DT, q[i], p1)
    assert r["iters"] >= 0
    acc(tot, r["counts"]); its += r["iters"]
out["dual_pendulums_step"] = {"unit": "DEL step", "sample": "%d rollouts x %d steps of W4" % (n, steps),
                              "flops": OC.flops(tot) / (n * steps), "sincos": tot["sincos"] / (n * steps),
                              "newton_iters": its / (n * steps)}

# ---- pend-on-cart linearization: random states (bench.py's first secondary line) and exact hints (W3)
d = systems.named_desc("pend_on_cart1")
n = 512
q1 = rng.uniform(-0.5, 0.5, (n, 2)); p1 = rng.normal(0, 1, (n, 2)); u1 = rng.uniform(-2, 2, (n, 1))
tot, its, tot0 = {}, 0, {}
for i in range(n):
    r = OC.linearize(d, 0.0, DT, q1[i], p1[i], u1[i], np.zeros(0))
    assert r["iters"] >= 0
    acc(tot, r["counts"]); its += r["iters"]
    q2 = H.step(d, 1, 0.0, DT, q1[i], p1[i], u1=u1[i][None, :])[1]
    r0 = OC.linearize(d, 0.0, DT, q1[i], p1[i], u1[i], np.zeros(0), q2_guess=q2[:d.nd])
    assert r0["iters"] == 0
    acc(tot0, r0["counts"])
out["pend_on_cart1_linearization"] = {"unit": "linearization", "sample": "%d random states" % n, "flops": OC.flops(tot) / n,
                                      "sincos": tot["sincos"] / n, "newton_iters": its / n}
out["pend_on_cart1_linearization_exact_hint"] = {"unit": "linearization", "sample": "%d states, hint = the step's solution (W3)" % n,
                                                 "flops": OC.flops(tot0) / n, "sincos": tot0["sincos"] / n, "newton_iters": 0.0}

# ---- W5: marionette linearization at perturbed points of the recorded rollout (bench.py's W5 inputs)
d = systems.named_desc("puppet")
g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
n = 24
idx = rng.integers(1, 58, n)
q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
q1[:, :d.nd] += rng.normal(0, 0.02, (n, d.nd)); p1 += rng.normal(0, 0.02, (n, d.nd))
tot, its = {}, 0
for i in range(n):
    r = OC.linearize(d, 0.0, DT, q1[i], p1[i], np.zeros(0), g["roll_k2"][idx[i]], lam_guess=g["roll_lambda"][idx[i] - 1])
    assert r["iters"] >= 0, r["iters"]
    acc(tot, r["counts"]); its += r["iters"]
out["puppet_linearization"] = {"unit": "linearization", "sample": "%d perturbed points of the recorded rollout (W5)" % n,
                               "flops": OC.flops(tot) / n, "sincos": tot["sincos"] / n, "newton_iters": its / n,
                               "note": "frame formulation of the thread-per-instance path (86 frames); the cooperative kernels' link "
                                       "formulation executes 4.98e5 flops for the same result (profiles/flops.json)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "opcounts.json"), "w"), indent=1)
for k, v in out.items():
    if isinstance(v, dict):
        print("%-42s %12.1f flops + %6.1f sin/cos per %s  (%.2f Newton iterations)" % (k, v["flops"], v["sincos"], v["unit"], v["newton_iters"]))
