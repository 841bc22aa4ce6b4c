#!/usr/bin/env python3
"""Generate golden vectors from the reference itself (``oracle/_ref``, built by
``oracle/build_ref.py``) and commit them as small fixtures under ``tests/golden/``.

TEST INFRASTRUCTURE ONLY.

For every named system of BASELINE.json's configs this script
  * builds the system with the *reference's* own Python API exactly as the cited example does,
  * flattens it with ``trep_b200.model.flatten_trep_system`` and checks that the native model
    mirror (``trep_b200.systems``) yields the identical description (the marionette's
    description is written to ``trep_b200/data/puppet.json``),
  * runs the reference ``MidpointVI`` on seeded inputs and records, per case:
      inputs  t1,t2,q1,p1,u1,k2,q2_guess,lambda_guess
      outputs q2,p2,lambda1,iters, every deriv1 array, A/B of DSystem.fdx/fdu,
              and (where the reference supports it) every deriv2 tensor,
  * records multi-step rollouts (final state, sampled states, Newton-iteration histogram).

Array layouts are the reference's raw storage layouts (``trep/midpointvi.py:40-136``):
first-derivative arrays are [wrt-index][output-index]; second-derivative arrays
[wrtA][wrtB][output].

Usage:  python oracle/gen_golden.py            (writes tests/golden/*.npz, trep_b200/data/puppet.json)
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_systems as R  # noqa: E402

trep = R.trep
discopt = R.discopt
REF_BUILDERS = R.REF_BUILDERS
ref_tase_pendulum, ref_puppet = R.ref_tase_pendulum, R.ref_puppet

from trep_b200 import model as M  # noqa: E402
from trep_b200 import systems as S  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "trep_b200", "data")

D1_NAMES = ["q2_dq1", "q2_dp1", "q2_du1", "q2_dk2", "p2_dq1", "p2_dp1", "p2_du1", "p2_dk2",
            "l1_dq1", "l1_dp1", "l1_du1", "l1_dk2"]
D2_SUFFIX = ["dq1dq1", "dq1dp1", "dq1du1", "dq1dk2", "dp1dp1", "dp1du1", "dp1dk2", "du1du1",
             "du1dk2", "dk2dk2"]
D2_NAMES = [p + "_" + s for p in ("q2", "p2", "l1") for s in D2_SUFFIX]


# ---- recording ---------------------------------------------------------------------------------
def record_case(mvi, dsys_time, t1, t2, q1, p1, u1, k2, q2_guess, lam_guess, want_d2):
    """One `DSystem.set`-style evaluation: initialize_from_state + step (+ derivs)."""
    nd = mvi.nd
    mvi.initialize_from_state(t1, q1, p1)
    q2_hint = None if q2_guess is None else np.array(q2_guess)
    iters = mvi.step(t2, u1, k2, q2_hint=q2_hint, lambda1_hint=lam_guess)
    out = dict(t1=t1, t2=t2, q1=np.array(q1, float), p1=np.array(p1, float), u1=np.array(u1, float),
               k2=np.array(k2, float),
               q2_guess=np.array(q1[:nd] if q2_guess is None else q2_guess[:nd], float),
               lambda_guess=np.zeros(mvi.nc) if lam_guess is None else np.array(lam_guess, float),
               q2=mvi.q2, p2=mvi.p2, lambda1=mvi.lambda1, iters=np.int32(iters))
    mvi._calc_deriv1()
    for n in D1_NAMES:
        out[n] = np.array(getattr(mvi, "_" + n))
    # A/B exactly as DSystem.fdx/fdu assembles them (trep/discopt/dsystem.py:284-317)
    dsys = discopt.DSystem(mvi, dsys_time)
    dsys._k = 0
    dsys._time = np.array([t1, t2])
    out["A"] = dsys.fdx()
    out["B"] = dsys.fdu()
    if want_d2:
        mvi._calc_deriv2()
        for n in D2_NAMES:
            out[n] = np.array(getattr(mvi, "_" + n))
    return out


def stack(cases):
    keys = cases[0].keys()
    return {"case_" + k: np.stack([np.asarray(c[k]) for c in cases]) for k in keys}


def rollout(mvi, q0, q1, dt, nsteps, u_fn=None, k_fn=None, sample=50):
    """initialize_from_configs + nsteps of step(); records sampled states + iteration counts."""
    mvi.initialize_from_configs(0.0, q0, dt, q1)
    p0 = mvi.p2
    qs, ps, its, ls = [mvi.q2], [mvi.p2], [], []
    for s in range(nsteps):
        u1 = tuple() if u_fn is None else u_fn(mvi.t2)
        k2 = tuple() if k_fn is None else k_fn(mvi.t2)
        its.append(mvi.step(mvi.t2 + dt, u1, k2))
        if (s + 1) % sample == 0 or s == nsteps - 1:
            qs.append(mvi.q2); ps.append(mvi.p2); ls.append(mvi.lambda1)
    return dict(roll_q0=np.array(q0, float), roll_q1=np.array(q1, float), roll_dt=dt,
                roll_nsteps=np.int32(nsteps), roll_sample=np.int32(sample), roll_p_init=p0,
                roll_q=np.array(qs), roll_p=np.array(ps), roll_lambda=np.array(ls),
                roll_iters=np.array(its, np.int32))


def gen_small(name, rng, ncases, dt, want_d2, q_lo, q_hi, p_scale, u_scale):
    system = REF_BUILDERS[name]()
    mvi = trep.MidpointVI(system, num_threads=1)
    nq, nd, nu = mvi.nq, mvi.nd, mvi.nu
    cases = []
    for c in range(ncases):
        q1 = rng.uniform(q_lo, q_hi, nq)
        p1 = rng.normal(0.0, p_scale, nd)
        u1 = rng.uniform(-u_scale, u_scale, nu)
        t1 = 0.01 * c
        # half the cases start Newton from q1 (reference default), half from a nearby hint
        hint = None if c % 2 == 0 else q1[:nd] + rng.normal(0, 1e-2, nd)
        cases.append(record_case(mvi, None, t1, t1 + dt, q1, p1, u1, np.zeros(0), hint, None, want_d2))
    return system, mvi, cases


def main():
    os.makedirs(GOLD, exist_ok=True)
    os.makedirs(DATA, exist_ok=True)
    rng = np.random.default_rng(0)

    # -- descriptions: reference flattening must equal the native model mirror ------------------
    descs = {}
    for name in S.NAMED:
        ref_sys = REF_BUILDERS[name]()
        d = M.flatten_trep_system(ref_sys, name=name)
        descs[name] = d
        if name == "puppet":
            d.save(os.path.join(DATA, "puppet.json"))
            np.save(os.path.join(GOLD, "puppet_q0.npy"), np.array(ref_sys.q))
        else:
            native = S.named_desc(name)
            assert native.equal(d), "native model mirror disagrees with the reference for " + name
        print("desc", name, "frames", d.n_frames, "nd", d.nd, "nk", d.nk, "nu", d.nu, "nc", d.nc)

    # -- known-answer test of examples/papers/tase2012/pend-single-step.py ---------------------
    system = ref_tase_pendulum()
    mvi = trep.MidpointVI(system, num_threads=1)
    cases = [record_case(mvi, None, 0.0, 0.1, [0.2], [0.5], [0.8], np.zeros(0), None, None, True)]
    for c in range(7):
        q1 = rng.uniform(-math.pi, math.pi, 1)
        cases.append(record_case(mvi, None, 0.0, 0.1, q1, rng.normal(0, 1, 1), rng.uniform(-1, 1, 1),
                                 np.zeros(0), None, None, True))
    np.savez_compressed(os.path.join(GOLD, "tase_pendulum.npz"), **stack(cases))

    # -- small unconstrained systems ----------------------------------------------------------
    spec = {
        #  name            ncases dt   d2    q_lo      q_hi     p_scale u_scale
        "pendulum1":       (16, 0.01, True, -math.pi, math.pi, 2.0, 0.0),
        "pendulum5":       (8, 0.01, True, -math.pi, math.pi, 2.0, 0.0),
        "damped_pendulum": (16, 0.01, True, -math.pi, math.pi, 5.0, 0.0),
        "pend_on_cart1":   (16, 0.01, True, -math.pi, math.pi, 3.0, 2.0),
        "pend_on_cart2":   (16, 0.01, True, -math.pi, math.pi, 3.0, 2.0),
        # LinearSpring has no C V_dqdqdq (trep/_trep/potentials/linearspring.c) -> no deriv2
        "dual_pendulums":  (16, 0.01, False, -math.pi, math.pi, 3.0, 0.0),
    }
    for name, (ncases, dt, d2, lo, hi, ps, us) in spec.items():
        system, mvi, cases = gen_small(name, rng, ncases, dt, d2, lo, hi, ps, us)
        out = stack(cases)
        # rollouts
        nq = mvi.nq
        if name == "damped_pendulum":
            out.update(rollout(mvi, (0.23,), (0.24,), 0.01, 1001, sample=100))
        elif name == "dual_pendulums":
            out.update(rollout(mvi, [3, -3], [3, -3], 0.01, 1000, sample=100))
        elif name.startswith("pendulum"):
            q0 = np.zeros(nq); q0[0] = math.pi / 4
            out.update(rollout(mvi, q0, q0, 0.01, 1000, sample=100))
        elif name.startswith("pend_on_cart"):
            nu = mvi.nu
            out.update(rollout(mvi, [0.0, 0.3], [0.0, 0.3], 0.01, 1000,
                               u_fn=lambda t: tuple([1.5 * math.sin(2.0 * t)] + [0.2 * math.cos(t)] * (nu - 1)),
                               sample=100))
            out["roll_u_desc"] = np.array("u0=1.5*sin(2 t2_prev); u1=0.2*cos(t2_prev)")
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print("golden", name, {k: v.shape for k, v in out.items() if k in ("case_q2", "case_A", "roll_q")})

    # -- marionette ---------------------------------------------------------------------------
    puppet = ref_puppet()
    mvi = trep.MidpointVI(puppet, num_threads=1)
    q0 = puppet.q
    qk0 = np.array(puppet.qk)
    dt = 0.01
    idx = {n: puppet.get_config(n + '_string-length').k_index
           for n in ("left_leg", "right_leg", "left_arm", "right_arm")}

    def k_fn(t):
        qk = qk0.copy()
        s = 0.1 * math.sin(0.6 * math.pi * t)
        qk[idx["left_leg"]] -= s; qk[idx["right_leg"]] += s
        qk[idx["left_arm"]] += s; qk[idx["right_arm"]] -= s
        return tuple(qk)

    nsteps = 60
    mvi.initialize_from_configs(0.0, q0, dt, q0)
    traj = [dict(q=mvi.q2, p=mvi.p2, lam=mvi.lambda1, t=mvi.t2)]
    its = []
    for s in range(nsteps):
        k2 = k_fn(mvi.t2)
        its.append(mvi.step(mvi.t2 + dt, tuple(), k2))
        traj.append(dict(q=mvi.q2, p=mvi.p2, lam=mvi.lambda1, t=mvi.t2, k2=np.array(k2)))
    out = dict(roll_q0=np.array(q0), roll_q1=np.array(q0), roll_dt=dt, roll_nsteps=np.int32(nsteps),
               roll_sample=np.int32(1), roll_q=np.array([x["q"] for x in traj]),
               roll_p=np.array([x["p"] for x in traj]), roll_lambda=np.array([x["lam"] for x in traj[1:]]),
               roll_k2=np.array([x["k2"] for x in traj[1:]]), roll_iters=np.array(its, np.int32))
    # cases: points of the trajectory (warm-started exactly like mvi.step does), plus perturbed ones
    nd = mvi.nd
    cases = []
    pick = [1, 7, 19, 33, 48, 59]
    for j, s in enumerate(pick):
        a, b = traj[s], traj[s + 1]
        q1 = a["q"].copy(); p1 = a["p"].copy()
        if j >= 3:  # perturbed state (inconsistent trajectory point -> Newton has real work)
            q1[:nd] += rng.normal(0, 0.05, nd)
            p1 += rng.normal(0, 0.05, nd)
        cases.append(record_case(mvi, None, a["t"], b["t"], q1, p1, np.zeros(0), b["k2"],
                                 None, a["lam"], want_d2=(j in (0, 3))))
    # deriv2 only on two cases (1.76 MB each): store those separately
    d2cases = [{n: c[n] for n in D2_NAMES} for c in cases if "q2_dq1dq1" in c]
    for c in cases:
        for n in D2_NAMES:
            c.pop(n, None)
    out.update(stack(cases))
    np.savez_compressed(os.path.join(GOLD, "puppet.npz"), **out)
    # second-derivative goldens for the marionette: keep float64 but only the first case in full
    d2 = {"case_index": np.array([0, 3], np.int32)}
    for n in D2_NAMES:
        d2["case_" + n] = np.stack([c[n] for c in d2cases[:1]])
    np.savez_compressed(os.path.join(GOLD, "puppet_deriv2.npz"), **d2)
    print("golden puppet: iters", its[:10], "...", "case iters", [int(c["iters"]) for c in cases])


if __name__ == "__main__":
    main()
