"""Golden vectors for the plugin kinds beyond BASELINE.json's configs:
PointOnPlane constraints (examples/pccd.py) and Body / Hybrid / Spatial wrenches.

TEST INFRASTRUCTURE (the oracle side): runs the reference itself (oracle/_ref) and records
tests/golden/{pccd,wrench_arm}.npz in the layout of oracle/gen_golden.py.  Kept separate so that
the fixtures of the BASELINE systems are not regenerated.  Usage:  python oracle/gen_golden_f4.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import gen_golden as GG  # noqa: E402  (record_case / stack / rollout helpers)
from ref_systems import trep, REF_BUILDERS  # noqa: E402
from trep_b200 import model as M  # noqa: E402
from trep_b200 import systems as S  # noqa: E402


def main():
    rng = np.random.default_rng(4)
    # spline tables fitted by the reference's trep.Spline (host side, trep/spline.py) for the mirror
    import json
    sp = trep.Spline(S.SPLINE_DATA)
    with open(os.path.join(ROOT, "trep_b200", "data", "spline_pendulum.json"), "w") as fh:
        json.dump({"x_points": np.array(sp._x_points).tolist(),
                   "coefficients": np.array(sp._coefficients).tolist()}, fh)
    for name in S.EXTRA:
        d = M.flatten_trep_system(REF_BUILDERS[name](), name=name)
        assert S.named_desc(name).equal(d), "native model mirror disagrees with the reference for " + name
        print("desc", name, "frames", d.n_frames, "nd", d.nd, "nk", d.nk, "nu", d.nu, "nc", d.nc)

    # ---- pccd: consistent states along a reference rollout, half of them perturbed -------------
    system = REF_BUILDERS["pccd"]()
    mvi = trep.MidpointVI(system, num_threads=1)
    q0 = system.q
    dt = 0.01
    mvi.initialize_from_configs(0.0, q0, dt, q0)
    traj = [dict(q=mvi.q2, p=mvi.p2, lam=mvi.lambda1, t=mvi.t2)]
    its = []
    nsteps = 300
    for s in range(nsteps):
        its.append(mvi.step(mvi.t2 + dt))
        traj.append(dict(q=mvi.q2, p=mvi.p2, lam=mvi.lambda1, t=mvi.t2))
    out = dict(roll_q0=np.array(q0), roll_q1=np.array(q0), roll_dt=dt, roll_nsteps=np.int32(nsteps),
               roll_sample=np.int32(1), roll_q=np.array([x["q"] for x in traj]),
               roll_p=np.array([x["p"] for x in traj]), roll_lambda=np.array([x["lam"] for x in traj[1:]]),
               roll_iters=np.array(its, np.int32))
    nd = mvi.nd
    cases = []
    for j, s in enumerate([1, 40, 90, 150, 210, 260, 280, 299]):
        a, b = traj[s], traj[s + 1]
        q1 = a["q"].copy(); p1 = a["p"].copy()
        if j % 2 == 1:
            q1[:nd] += rng.normal(0, 0.02, nd)
            p1 += rng.normal(0, 0.05, nd)
        cases.append(GG.record_case(mvi, None, a["t"], b["t"], q1, p1, np.zeros(0), np.zeros(0),
                                    None, a["lam"], want_d2=True))
    out.update(GG.stack(cases))
    np.savez_compressed(os.path.join(GG.GOLD, "pccd.npz"), **out)
    print("golden pccd: rollout iters", sorted(set(its)), "case iters", [int(c["iters"]) for c in cases])

    # ---- wrench arm: random states and inputs --------------------------------------------------
    system = REF_BUILDERS["wrench_arm"]()
    mvi = trep.MidpointVI(system, num_threads=1)
    nq, nu = mvi.nq, mvi.nu
    cases = []
    for c in range(12):
        q1 = np.array([rng.uniform(-math.pi, math.pi), rng.uniform(-1.2, 1.2), rng.uniform(-0.3, 0.5)])
        p1 = rng.normal(0, 1.0, nq)
        u1 = rng.uniform(-2, 2, nu)
        hint = None if c % 2 == 0 else q1 + rng.normal(0, 1e-2, nq)
        cases.append(GG.record_case(mvi, None, 0.02 * c, 0.02 * c + 0.01, q1, p1, u1, np.zeros(0), hint, None,
                                    want_d2=True))
    out = GG.stack(cases)
    out.update(GG.rollout(mvi, [0.3, -0.4, 0.1], [0.3, -0.4, 0.1], 0.01, 400,
                          u_fn=lambda t: (1.0 * math.sin(3 * t), 0.5 * math.cos(2 * t), 0.8 * math.sin(t), 0.3 * math.cos(t), -0.6 * math.sin(2 * t)),
                          sample=50))
    out["roll_u_desc"] = np.array("u = (sin(3 t), 0.5 cos(2 t), 0.8 sin(t), 0.3 cos(t), -0.6 sin(2 t)), t = t2 of the previous step")
    np.savez_compressed(os.path.join(GG.GOLD, "wrench_arm.npz"), **out)
    print("golden wrench_arm: case iters", [int(c["iters"]) for c in cases])
    spline_pendulum(np.random.default_rng(5))


def spline_pendulum(rng):
    # NonlinearConfigSpring: the reference's V_dqdqdq has the wrong sign
    # (nonlinear_config_spring.c:53: "-Spline_ddy * -m * m"), so its second-derivative tensors are not
    # recorded; the tests check those by finite differences of the first derivatives instead.
    system = REF_BUILDERS["spline_pendulum"]()
    mvi = trep.MidpointVI(system, num_threads=1)
    nq = mvi.nq
    cases = []
    for c in range(16):
        # m q + b sweeps every spline segment, including both extrapolation pieces
        q1 = np.array([rng.uniform(-2.4, 2.2), rng.uniform(-math.pi, math.pi)])
        p1 = rng.normal(0, 1.0, nq)
        hint = None if c % 2 == 0 else q1 + rng.normal(0, 1e-2, nq)
        cases.append(GG.record_case(mvi, None, 0.01 * c, 0.01 * c + 0.01, q1, p1, np.zeros(0), np.zeros(0), hint,
                                    None, want_d2=False))
    out = GG.stack(cases)
    out.update(GG.rollout(mvi, [1.2, -0.5], [1.2, -0.5], 0.01, 600, sample=100))
    np.savez_compressed(os.path.join(GG.GOLD, "spline_pendulum.npz"), **out)
    print("golden spline_pendulum: case iters", [int(c["iters"]) for c in cases])


if __name__ == "__main__":
    if "--spline-only" in sys.argv:
        spline_pendulum(np.random.default_rng(5))
        sys.exit(0)
    main()
