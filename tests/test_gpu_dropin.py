"""The drop-in, proven on the reference's own classes: trep.discopt.DSystem / DOptimizer of oracle/_ref (the
unmodified reference) against subclasses of those same classes whose loops over the time steps run on libtrepb.so
(trep_b200/binding.py = the binding INTEGRATION.md section 3 describes).  Same problem scripts as the reference's
examples; everything outside the replaced loops (cost, monitors, step / optimize logic) is the reference's code."""
import math
import os
import sys

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_systems as R
    return R


@pytest.fixture(scope="module")
def classes(ref):
    from trep_b200 import binding, lib
    assert lib.device_count() > 0
    D = binding.batched_dsystem_class(ref.discopt.DSystem, ref.trep.ConvergenceError)
    O = binding.batched_doptimizer_class(ref.discopt.DOptimizer, ref.trep.ConvergenceError)
    return D, O


def pend_on_cart_problem(ref, DSys, Opt, K=400):
    """examples/pend-on-cart-optimization.py:48-110 (the torque-input system, its initial trajectory and cost)."""
    trep, discopt = ref.trep, ref.discopt
    system = ref.ref_pend_on_cart(True)
    mvi = trep.MidpointVI(system, num_threads=1)
    t = np.arange(0.0, 0.01 * (K + 1) - 1e-9, 0.01)
    dsys = DSys(mvi, t)
    (X, U) = dsys.build_trajectory()
    for k in range(dsys.kf()):
        if k == 0:
            dsys.set(X[k], U[k], 0)
        else:
            dsys.step(U[k])
        X[k + 1] = dsys.f()
    qd = np.zeros((len(t), system.nQ))
    th = system.get_config('theta').index
    for i, ti in enumerate(t):
        if 1.0 <= ti <= 3.0:
            qd[i, th] = (1 - math.cos(2 * math.pi / 2 * (ti - 1.0))) * (130 * math.pi / 180) / 2
    (Xd, Ud) = dsys.build_trajectory(qd)
    wx = 0.01 * np.ones(dsys.nX); wx[system.get_config('x').index] = 0.01; wx[th] = 100.0
    wu = 0.01 * np.ones(dsys.nU)
    cost = discopt.DCost(Xd, Ud, np.diag(wx), np.diag(wu))
    return dsys, Opt(dsys, cost), X, U


@pytest.mark.parametrize("method", ["quasi", "newton"])
def test_one_doptimizer_step_of_pend_on_cart(ref, classes, method):
    """DOptimizer.step (doptimizer.py:466-505): descent direction (linearize_trajectory, projection gains, the
    newton model's second derivatives, the LQ solve), then the Armijo line search."""
    D, O = classes
    dsys0, opt0, X, U = pend_on_cart_problem(ref, ref.discopt.DSystem, ref.discopt.DOptimizer)
    dsys1, opt1, X1, U1 = pend_on_cart_problem(ref, D, O)
    assert np.array_equal(X, X1)
    # a first quasi step with the stock classes gives a trajectory with non-trivial inputs to start from
    X, U = opt0.step(0, X, U, "quasi")[1:3]
    d0 = opt0.calc_descent_direction(X, U, method)
    d1 = opt1.calc_descent_direction(X, U, method)
    G.assert_close(d1.Kproj, np.array(d0.Kproj), "Kproj", rtol=1e-8)
    G.assert_close(d1.dX, d0.dX, "dX", rtol=1e-7)
    G.assert_close(d1.dU, d0.dU, "dU", rtol=1e-7)
    n = len(X) - 1
    for name, f0, f1, cnt in (("Q", d0.Q, d1.Q, n + 1), ("R", d0.R, d1.R, n), ("S", d0.S, d1.S, n)):
        a = np.stack([f1(k) for k in range(cnt)]); b = np.stack([f0(k) for k in range(cnt)])
        G.assert_close(a, b, "newton model " + name, rtol=1e-9)
    s0 = opt0.step(1, X, U, method)
    s1 = opt1.step(1, X, U, method)
    assert s0.done == s1.done
    assert abs(s1.dcost0 - s0.dcost0) <= 1e-7 * abs(s0.dcost0)
    assert abs(s1.cost1 - s0.cost1) <= 1e-8 * abs(s0.cost1)
    assert s1.cost1 < opt0.calc_cost(X, U)
    G.assert_close(s1.nX, s0.nX, "new X", rtol=1e-6)
    G.assert_close(s1.nU, s0.nU, "new U", rtol=1e-6)


def test_linearize_and_project_of_the_reference_dsystem(ref, classes):
    D, O = classes
    dsys0, opt0, X, U = pend_on_cart_problem(ref, ref.discopt.DSystem, ref.discopt.DOptimizer, K=150)
    dsys1, opt1, _, _ = pend_on_cart_problem(ref, D, O, K=150)
    rng = np.random.default_rng(3)
    bX = X + rng.normal(0, 1e-3, X.shape); bU = U + rng.normal(0, 1e-2, U.shape)
    A0, B0 = dsys0.linearize_trajectory(bX, bU)
    A1, B1 = dsys1.linearize_trajectory(bX, bU)
    G.assert_close(A1, A0, "A")
    G.assert_close(B1, B0, "B")
    K0 = dsys0.calc_feedback_controller(bX, bU)
    K1 = dsys1.calc_feedback_controller(bX, bU)
    G.assert_close(K1, np.array(K0), "Kproj", rtol=1e-8)
    p0 = dsys0.project(bX, bU, K0)
    p1 = dsys1.project(bX, bU, K0)
    G.assert_close(p1.X, p0.X, "projected X", rtol=1e-9)
    G.assert_close(p1.U, p0.U, "projected U", rtol=1e-9)


def test_newton_model_of_a_marionette_trajectory(ref, classes):
    """calc_newton_model (doptimizer.py:319-345) on a 20-step trajectory of the marionette: adjoint recursion,
    then the z-contracted second derivatives of all 20 steps in ONE launch, against the reference's per-step
    set() + fdxdx(z) / fdxdu(z) / fdudu(z)."""
    D, O = classes
    g = G.golden("puppet")
    K = 20
    out = []
    for DSys, Opt in ((ref.discopt.DSystem, ref.discopt.DOptimizer), (D, O)):
        puppet = ref.ref_puppet()
        mvi = ref.trep.MidpointVI(puppet, num_threads=4)
        t = 0.01 * (1 + np.arange(K + 1))
        dsys = DSys(mvi, t)
        nd = mvi.nd
        Q = g["roll_q"][:K + 1]; P = g["roll_p"][:K + 1]
        V = np.zeros((K + 1, mvi.nk)); V[1:] = (Q[1:, nd:] - Q[:-1, nd:]) / 0.01
        X, U = dsys.build_trajectory(Q, P, V, None, g["roll_k2"][:K])
        Xd = X.copy(); Xd[:, :nd] += 0.05
        cost = ref.discopt.DCost(Xd, U.copy(), np.eye(dsys.nX), 0.1 * np.eye(dsys.nU))
        opt = Opt(dsys, cost)
        Kp, A, B = dsys.calc_feedback_controller(X, U, opt.Qproj, opt.Rproj, True)
        m = opt.calc_newton_model(X, U, A, B, Kp)
        out.append((np.array(Kp), A, B, np.stack([m.Q(k) for k in range(K + 1)]), np.stack([m.R(k) for k in range(K)]),
                    np.stack([m.S(k) for k in range(K)])))
    for name, a, b in zip(("Kproj", "A", "B", "Q", "R", "S"), out[1], out[0]):
        G.assert_close(a, b, "marionette " + name, rtol=1e-8 if name == "Kproj" else 1e-9)
