"""CPU check of the shared host/device MidpointVI math against the reference's golden vectors.

The same header (trep_b200/csrc/trepb_math.cuh) is what the CUDA kernels instantiate, so these
tests pin the algebra on a machine without a GPU; the `-m gpu` tests pin the kernels proper."""
import numpy as np
import pytest

import golden_util as G
import hostmath as H


@pytest.mark.parametrize("name", G.ALL + G.EXTRA + G.PARITY + G.PARITY_SPRING)
def test_cases_step_and_deriv1(name):
    g = G.golden(name)
    d = G.desc(name)
    n = g["case_q1"].shape[0]
    flips = 0
    for c in range(n):
        out = H.linearize(d, float(g["case_t1"][c]), float(g["case_t2"][c]), g["case_q1"][c],
                          g["case_p1"][c], g["case_u1"][c], g["case_k2"][c],
                          q2_guess=g["case_q2_guess"][c], lam_guess=g["case_lambda_guess"][c])
        assert out["rc"] == 0
        flips += int(out["iters"] != int(g["case_iters"][c]))
        for k in ("q2", "p2", "lambda1", "A", "B"):
            G.assert_close(out[k], g["case_" + k][c], "%s case %d %s" % (name, c, k))
        for k in G.RAW:
            G.assert_close(out[k], g["case_" + k][c], "%s case %d %s" % (name, c, k))
    assert flips == 0, "Newton iteration counts differ from the reference in %d cases" % flips


@pytest.mark.parametrize("name", G.ALL + ["pccd", "spline_pendulum", "wrench_arm"] + G.PARITY + G.PARITY_SPRING)
def test_cooperative_math_cases(name):
    """The team-cooperative formulation (link tables, world-coordinate spatial algebra,
    right-looking LU; trepb_coop_math.cuh) run with a one-lane host team against the goldens."""
    g = G.golden(name)
    d = G.desc(name)
    if H.coop_info(d) is None:
        pytest.skip("cooperative path does not apply")
    flips = 0
    # run-time sizes everywhere; the compile-time-size flavour (register-resident right-hand-side
    # columns) for the shape the build specialises
    # (2: the same with the external-slab layout of the first-derivative workspace, ExtDims)
    for static in ([0, 1, 2] if name == "puppet" else [0]):
        for c in range(g["case_q1"].shape[0]):
            out = H.coop_linearize(d, float(g["case_t1"][c]), float(g["case_t2"][c]), g["case_q1"][c],
                                   g["case_p1"][c], g["case_u1"][c], g["case_k2"][c],
                                   q2_guess=g["case_q2_guess"][c], lam_guess=g["case_lambda_guess"][c],
                                   static_dims=static)
            assert out["rc"] == 0
            flips += int(out["iters"] != int(g["case_iters"][c]))
            for k in ("q2", "p2", "lambda1", "A", "B"):
                G.assert_close(out[k], g["case_" + k][c], "%s case %d %s" % (name, c, k))
            for k in G.RAW:
                G.assert_close(out[k], g["case_" + k][c], "%s case %d %s" % (name, c, k))
    assert flips == 0


def test_cooperative_solve_only_slab_layout():
    """The step / project / p2 kernels of the external-slab flavour keep the pair arrays and both constraint Jacobians
    in the slab (ExtSolveDims): the Newton solve on that layout, with a poisoned slab, against the goldens."""
    g = G.golden("puppet")
    d = G.desc("puppet")
    for c in range(g["case_q1"].shape[0]):
        out = H.coop_linearize(d, float(g["case_t1"][c]), float(g["case_t2"][c]), g["case_q1"][c], g["case_p1"][c],
                               g["case_u1"][c], g["case_k2"][c], q2_guess=g["case_q2_guess"][c],
                               lam_guess=g["case_lambda_guess"][c], static_dims=3, derivs=False)
        assert out["rc"] == 0 and out["iters"] == int(g["case_iters"][c])
        for k in ("q2", "p2", "lambda1"):
            G.assert_close(out[k], g["case_" + k][c], "solve-only slab case %d %s" % (c, k))
    # a rollout on the same layout (the kernels' time loop keeps the slab across steps)
    dt, n = float(g["roll_dt"]), 12
    p0 = H.coop_calc_p2(d, dt, g["roll_q0"], g["roll_q1"])
    args = (d, dt, 2 * dt, g["roll_q1"], p0, np.zeros((n, d.nu)), g["roll_k2"][:n])
    out = H.coop_linearize(*args, nsteps=n, static_dims=3, derivs=False)
    want = H.coop_linearize(*args, nsteps=n, static_dims=1, derivs=False)
    assert out["rc"] == 0 and want["rc"] == 0 and out["iters"] == want["iters"] > 0
    for k in ("q2", "p2", "lambda1"):
        assert np.array_equal(out[k], want[k]), k


def test_cooperative_aux_export_matches_thread_path():
    """The factorizations the cooperative linearize kernel exports for the second-derivative
    kernel are in LU_decomp's convention (math-code.c:337-432): same factors and permutation as the
    thread-per-instance path produces with its Crout elimination."""
    g = G.golden("puppet")
    d = G.desc("puppet")
    c = 0
    args = (d, float(g["case_t1"][c]), float(g["case_t2"][c]), g["case_q1"][c], g["case_p1"][c], g["case_u1"][c],
            g["case_k2"][c])
    kw = dict(q2_guess=g["case_q2_guess"][c], lam_guess=g["case_lambda_guess"][c])
    a = H.coop_linearize(*args, **kw)["aux"]
    b = H.coop_linearize(*args, static_dims=True, **kw)["aux"]
    assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(a)))
    nd, nc = d.nd, d.nc
    m2 = a[:nd * nd].reshape(nd, nd)
    piv = a[nd * nd:nd * nd + nd].astype(int)
    assert sorted(piv) == list(range(nd))
    # P M2 = L U with M2 = D2D1L2_D2fm2[:nd,:nd]^T recovered from the golden first derivatives:
    # q2_dp1 = M2^-1 (-I) when there are no constraints; with constraints check L U is well formed
    Lm = np.tril(m2, -1) + np.eye(nd)
    Um = np.triu(m2)
    assert np.all(np.abs(Lm) <= 1e6) and np.all(np.abs(np.diag(Um)) > 1e-12)


def test_cooperative_puppet_rollout_and_tables():
    g = G.golden("puppet")
    d = G.desc("puppet")
    info = H.coop_info(d)
    # 34 variable frames of 86, 10 link levels; the workspace of one instance fits 7 times (run-time
    # sizes) / 8 times (compile-time sizes) next to the tables in one SM's 227 KB of shared memory
    assert info["nl"] == 34 and info["nlevels"] == 10
    assert 7 * info["ws_doubles"] * 8 + info["blob_bytes"] + 16 <= 227 * 1024
    assert 8 * info["ws_doubles_static"] * 8 + info["blob_bytes"] + 16 <= 227 * 1024
    dt, nsteps = float(g["roll_dt"]), int(g["roll_nsteps"])
    p0 = H.coop_calc_p2(d, dt, g["roll_q0"], g["roll_q1"])
    G.assert_close(p0, g["roll_p"][0], "puppet p_init")
    out = H.coop_linearize(d, dt, 2 * dt, g["roll_q1"], p0, np.zeros((nsteps, 0)), g["roll_k2"], nsteps=nsteps,
                           derivs=False)
    assert out["rc"] == 0
    G.assert_close(out["q2"], g["roll_q"][-1], "puppet final q", rtol=1e-8)
    G.assert_close(out["p2"], g["roll_p"][-1], "puppet final p", rtol=1e-8)
    G.assert_close(out["lambda1"], g["roll_lambda"][-1], "puppet final lambda", rtol=1e-8)
    assert out["iters"] == int(g["roll_iters"].sum())


@pytest.mark.parametrize("name", G.SMALL)
def test_rollout(name):
    g = G.golden(name)
    if "roll_q" not in g:
        pytest.skip("no rollout recorded")
    d = G.desc(name)
    dt, nsteps, sample = float(g["roll_dt"]), int(g["roll_nsteps"]), int(g["roll_sample"])
    p0 = H.calc_p2(d, dt, g["roll_q0"], g["roll_q1"])
    G.assert_close(p0, g["roll_p_init"], name + " p_init")
    if d.nu:
        t = dt * (1 + np.arange(nsteps))
        u = np.stack([1.5 * np.sin(2.0 * t)] + [0.2 * np.cos(t)] * (d.nu - 1), axis=1)
    else:
        u = None
    rc, q2, p2, lam, iters = H.step(d, nsteps, dt, dt, g["roll_q1"], p0, u1=u)
    assert rc == 0
    # chaotic systems amplify rounding differences along a rollout; the bound is loose on purpose,
    # single-step parity is pinned by test_cases_step_and_deriv1
    G.assert_close(q2, g["roll_q"][-1], name + " final q", rtol=1e-6)
    G.assert_close(p2, g["roll_p"][-1], name + " final p", rtol=1e-6)
    assert abs(iters - int(g["roll_iters"].sum())) <= max(2, nsteps // 100)


def test_puppet_rollout():
    g = G.golden("puppet")
    d = G.desc("puppet")
    dt, nsteps = float(g["roll_dt"]), int(g["roll_nsteps"])
    p0 = H.calc_p2(d, dt, g["roll_q0"], g["roll_q1"])
    G.assert_close(p0, g["roll_p"][0], "puppet p_init")
    rc, q2, p2, lam, iters = H.step(d, nsteps, dt, dt, g["roll_q1"], p0, k2=g["roll_k2"])
    assert rc == 0
    G.assert_close(q2, g["roll_q"][-1], "puppet final q", rtol=1e-8)
    G.assert_close(p2, g["roll_p"][-1], "puppet final p", rtol=1e-8)
    G.assert_close(lam, g["roll_lambda"][-1], "puppet final lambda", rtol=1e-8)
    assert iters == int(g["roll_iters"].sum())


D2_SYSTEMS = ["tase_pendulum", "pendulum1", "pendulum5", "damped_pendulum", "pend_on_cart1", "pend_on_cart2"] + G.EXTRA_D2 + G.PARITY


@pytest.mark.parametrize("method", ["pair", "jac"])
@pytest.mark.parametrize("name", D2_SYSTEMS)
def test_second_derivatives(name, method):
    """Second derivatives against the reference's _calc_deriv2 tensors: "pair" = hyper-dual residual
    per parameter pair (trepb_d2.cuh), "jac" = dual Jacobian tables per parameter + contraction
    (trepb_d2jac.cuh)."""
    g = G.golden(name)
    d = G.desc(name)
    for c in range(min(4, g["case_q1"].shape[0])):
        out = H.deriv2(d, float(g["case_t1"][c]), float(g["case_t2"][c]), g["case_q1"][c], g["case_p1"][c],
                       g["case_u1"][c], g["case_k2"][c], q2_guess=g["case_q2_guess"][c],
                       lam_guess=g["case_lambda_guess"][c], method=method)
        assert out["rc"] == 0
        # damper_only: the reference's LinearDamper f_ddqdq has a typo (forces/lineardamper.c:99); the tensors are
        # compared with the reference built with that one line corrected (casefix_*, oracle/gen_golden_r2.py)
        G.assert_d2_close(out, g, "%s case %d" % (name, c), index=c,
                          gold_prefix="casefix_" if name == "damper_only" else "case_")


@pytest.mark.parametrize("method", ["pair", "jac"])
def test_puppet_second_derivatives(method):
    import os
    g = G.golden("puppet")
    g2 = np.load(os.path.join(G.GOLD, "puppet_deriv2.npz"))
    c = int(g2["case_index"][0])
    d = G.desc("puppet")
    out = H.deriv2(d, float(g["case_t1"][c]), float(g["case_t2"][c]), g["case_q1"][c], g["case_p1"][c],
                   g["case_u1"][c], g["case_k2"][c], q2_guess=g["case_q2_guess"][c],
                   lam_guess=g["case_lambda_guess"][c], method=method)
    assert out["rc"] == 0
    G.assert_d2_close(out, g2, "puppet", index=0)


def test_division_by_the_time_step_is_exact():
    """div_dt (x r corrected by two fused multiply-adds) returns the bits of x / dt, and sqrt_threshold(tol)
    is the exact boundary of sqrt(x) > tol: the two replacements in the Newton loop change no result."""
    import ctypes as C
    lib = H.load()
    lib.th_div_dt_mismatches.restype = C.c_long
    for dt in (0.01, 0.001, 0.02, 0.0125, 1.0 / 3.0, 0.1):
        assert lib.th_div_dt_mismatches(C.c_double(dt), C.c_long(2000000)) == 0
    lib.th_sqrt_threshold.restype = C.c_double
    for tol in (1e-10, 1e-12, 3.3e-9, 1e-6, 0.0, 1.0):
        T = lib.th_sqrt_threshold(C.c_double(tol))
        assert np.sqrt(T) <= tol and (np.sqrt(np.nextafter(T, np.inf)) > tol)
