"""Golden vectors for the constraint / force kinds of SURVEY.md 8a rows a7 and a9 that no BASELINE config
exercises: PointToPoint1D/2D/3D (constraints/point.c:16-55), Distance with a FIXED length
(constraints/distance.c:16-136, config == NULL) and the LinearDamper's second derivatives
(forces/lineardamper.c:60-107).

TEST INFRASTRUCTURE (the oracle side): runs the reference itself (oracle/_ref) and records
tests/golden/{fourbar,loop3d,rod,damper_only}.npz in the layout of oracle/gen_golden.py (cases with every
deriv1 array, A / B, all 30 second-derivative tensors; a rollout with every step's state, multipliers,
Newton iteration count and the inputs / kinematic configs that drove it).
Usage:  python oracle/gen_golden_r2.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import gen_golden as GG  # noqa: E402  (record_case / stack helpers)
from ref_systems import trep, REF_BUILDERS  # noqa: E402
from trep_b200 import model as M  # noqa: E402
from trep_b200 import systems as S  # noqa: E402

# name -> (steps, u(t), k(t)); t = t2 of the previous step
DRIVE = {
    "fourbar": (300, lambda t: (0.8 * math.sin(2.0 * t),), None),
    "loop3d": (200, None, None),
    "rod": (300, None, lambda t: (0.15 * math.sin(3.0 * t),)),
    # LinearSprings between two arms + a distance constraint (cooperative kernels, round 2); the reference cannot
    # run _calc_deriv2 with a LinearSpring (no C V_dqdqdq), so no second-derivative tensors are recorded
    "spring_arms": (300, lambda t: (0.6 * math.sin(2.0 * t),), lambda t: (0.1 * math.sin(3.0 * t),)),
}


def constrained(name, rng, want_d2=True):
    system = REF_BUILDERS[name]()
    mvi = trep.MidpointVI(system, num_threads=1)
    nsteps, u_fn, k_fn = DRIVE[name]
    q0 = np.array(system.q)
    dt = 0.01
    mvi.initialize_from_configs(0.0, q0, dt, q0)
    traj = [dict(q=mvi.q2, p=mvi.p2, lam=mvi.lambda1, t=mvi.t2)]
    its = []
    for s in range(nsteps):
        u1 = tuple() if u_fn is None else u_fn(mvi.t2)
        k2 = tuple() if k_fn is None else k_fn(mvi.t2)
        its.append(mvi.step(mvi.t2 + dt, u1, k2))
        traj.append(dict(q=mvi.q2, p=mvi.p2, lam=mvi.lambda1, t=mvi.t2, u=np.array(u1, float), k2=np.array(k2, float)))
    out = dict(roll_q0=q0, roll_q1=q0, roll_dt=dt, roll_nsteps=np.int32(nsteps), roll_sample=np.int32(1),
               roll_q=np.array([x["q"] for x in traj]), roll_p=np.array([x["p"] for x in traj]),
               roll_lambda=np.array([x["lam"] for x in traj[1:]]),
               roll_u=np.array([x["u"] for x in traj[1:]]), roll_k2=np.array([x["k2"] for x in traj[1:]]),
               roll_iters=np.array(its, np.int32))
    nd = mvi.nd
    cases = []
    pick = np.linspace(1, nsteps - 2, 10).astype(int)
    for j, s in enumerate(pick):
        a, b = traj[s], traj[s + 1]
        q1 = a["q"].copy(); p1 = a["p"].copy()
        if j % 2 == 1:      # perturbed (inconsistent) state: Newton has real work, the constraint pulls back
            q1[:nd] += rng.normal(0, 0.02, nd)
            p1 += rng.normal(0, 0.05, nd)
        cases.append(GG.record_case(mvi, None, a["t"], b["t"], q1, p1, b["u"], b["k2"], None, a["lam"], want_d2=want_d2))
    out.update(GG.stack(cases))
    np.savez_compressed(os.path.join(GG.GOLD, name + ".npz"), **out)
    print("golden", name, "rollout iters", sorted(set(its)), "case iters", [int(c["iters"]) for c in cases])


def damper_only(rng):
    system = REF_BUILDERS["damper_only"]()
    mvi = trep.MidpointVI(system, num_threads=1)
    cases = []
    for c in range(12):
        q1 = rng.uniform(-math.pi, math.pi, 2)
        p1 = rng.normal(0, 3.0, 2)
        hint = None if c % 2 == 0 else q1 + rng.normal(0, 1e-2, 2)
        cases.append(GG.record_case(mvi, None, 0.01 * c, 0.01 * c + 0.01, q1, p1, np.zeros(0), np.zeros(0), hint, None,
                                    want_d2=True))
    out = GG.stack(cases)
    out.update(GG.rollout(mvi, [2.0, -1.0], [2.0, -1.0], 0.01, 500, sample=50))
    out.update(damper_only_fixed(out))
    np.savez_compressed(os.path.join(GG.GOLD, "damper_only.npz"), **out)
    print("golden damper_only: case iters", [int(c["iters"]) for c in cases])


FIXED_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {here!r}); sys.path.insert(0, {root!r})
import build_ref
build_ref.build("/root/reference", out={out!r}, fix_lineardamper=True)
sys.path.insert(0, {out!r})
import trep
from trep import tx, ty, tz, rx, ry, rz
system = trep.System()
system.import_frames([rx('theta1'), [tz(2, mass=1, name='pend1')], ty(1), [rx('theta2'), [tz(2, mass=1, name='pend2')]]])
trep.forces.LinearDamper(system, 'pend1', 'pend2', c=1)
trep.potentials.Gravity(system, name="Gravity")
mvi = trep.MidpointVI(system, num_threads=1)
g = np.load({inp!r})
names = {names!r}
res = {{n: [] for n in names}}
for c in range(g["q1"].shape[0]):
    mvi.initialize_from_state(float(g["t1"][c]), g["q1"][c], g["p1"][c])
    mvi.step(float(g["t2"][c]), tuple(), tuple(), q2_hint=np.array(g["q2_guess"][c]))
    mvi._calc_deriv1(); mvi._calc_deriv2()
    for n in names:
        res[n].append(np.array(getattr(mvi, "_" + n)))
np.savez({outp!r}, **{{n: np.stack(v) for n, v in res.items()}})
"""


def damper_only_fixed(out):
    """Second-derivative tensors of the same cases from the reference with forces/lineardamper.c:99 corrected
    (built in a temp dir by a child process; see build_ref.build): stored as casefix_*."""
    import subprocess
    import tempfile
    tmp = tempfile.mkdtemp(prefix="trep_ref_fixed_")
    inp, outp = os.path.join(tmp, "in.npz"), os.path.join(tmp, "out.npz")
    np.savez(inp, t1=out["case_t1"], t2=out["case_t2"], q1=out["case_q1"], p1=out["case_p1"], q2_guess=out["case_q2_guess"])
    script = FIXED_SCRIPT.format(here=HERE, root=ROOT, out=os.path.join(tmp, "ref"), inp=inp, outp=outp, names=GG.D2_NAMES)
    subprocess.check_call([sys.executable, "-c", script])
    r = np.load(outp)
    fixed = {"casefix_" + n: r[n] for n in GG.D2_NAMES}
    worst = max(float(np.max(np.abs(fixed["casefix_" + n] - out["case_" + n])) / max(1e-300, np.max(np.abs(out["case_" + n]))))
                for n in GG.D2_NAMES if out["case_" + n].size)
    print("damper_only: reference with lineardamper.c:99 corrected differs from the stock reference by up to %.3e "
          "(relative to each tensor's largest entry)" % worst)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return fixed


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "springs":      # only the round-2 spring fixture
        for name in S.PARITY_SPRING:
            d = M.flatten_trep_system(REF_BUILDERS[name](), name=name)
            assert S.named_desc(name).equal(d), "native model mirror disagrees with the reference for " + name
            print("desc", name, "frames", d.n_frames, "nd", d.nd, "nk", d.nk, "nu", d.nu, "nc", d.nc)
            constrained(name, np.random.default_rng(7), want_d2=False)
        return
    for name in S.PARITY:
        d = M.flatten_trep_system(REF_BUILDERS[name](), name=name)
        assert S.named_desc(name).equal(d), "native model mirror disagrees with the reference for " + name
        print("desc", name, "frames", d.n_frames, "nd", d.nd, "nk", d.nk, "nu", d.nu, "nc", d.nc)
    rng = np.random.default_rng(6)
    for name in ("fourbar", "loop3d", "rod"):
        constrained(name, rng)
    damper_only(rng)


if __name__ == "__main__":
    main()
