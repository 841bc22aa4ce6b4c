"""Plug-in kernels (trepb_load_plugin): a user system the library was not built with runs on ahead-of-time
specialised, register-resident kernels after its structure was compiled into a plug-in.  (Last test module on
purpose: loading the plug-in changes which kernel later System() calls for that structure pick.)"""
import os
import shutil

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from trep_b200 import lib as L
    if L.device_count() < 1:
        pytest.skip("needs a CUDA device")
    return L


def test_plugin_kernel_matches_the_reference(lib):
    from trep_b200 import build
    name = "damper_only"
    d = G.desc(name)
    g = G.golden(name)
    before = lib.System(d)
    assert not before.specialized and before.kernel_name == "general"
    ref = before.linearize(g["case_q1"], g["case_p1"], g["case_u1"], g["case_k2"], t1=g["case_t1"], t2=g["case_t2"],
                           q2_guess=g["case_q2_guess"], lambda_guess=g["case_lambda_guess"], want_raw=True)
    path = os.path.join(os.path.dirname(build.LIB), "libtrepb_plugin_%s.so" % name)
    if shutil.which(build.NVCC) is not None:
        path = build.build_plugin(d, name)          # rebuilt only when older than the headers / the library
    elif not os.path.exists(path):
        pytest.skip("plug-in not prebuilt and no nvcc on this box")
    assert lib.load_plugin(path) == 1
    s = lib.System(d)
    assert s.specialized and s.kernel_name == name
    out = s.linearize(g["case_q1"], g["case_p1"], g["case_u1"], g["case_k2"], t1=g["case_t1"], t2=g["case_t2"],
                      q2_guess=g["case_q2_guess"], lambda_guess=g["case_lambda_guess"], want_raw=True)
    assert np.all(out["status"] == 0)
    for k in ("q2", "p2", "A", "B"):
        G.assert_close(out[k], g["case_" + k], "plugin %s" % k)
        G.assert_close(out[k], ref[k], "plugin vs table-driven %s" % k)
    for k in G.RAW:
        G.assert_close(out[k], g["case_" + k], "plugin %s" % k)
    assert np.array_equal(out["iters"], g["case_iters"])
    # throughput: plug-in against the table-driven kernel on the same batch
    rng = np.random.default_rng(0)
    B = 1 << 18
    q = rng.uniform(-1.0, 1.0, (B, d.nq))
    p = rng.normal(0, 0.5, (B, d.nd))
    t = {}
    for label, sysm in (("general", before), ("plugin", s)):
        o = sysm.step(q, p, 0.01, 0.01, nsteps=20)
        t[label] = (o, sysm.last_kernel_ms())
    ok = (t["general"][0]["status"] == 0) & (t["plugin"][0]["status"] == 0)
    assert ok.mean() > 0.99
    G.assert_close(t["plugin"][0]["q2"][ok], t["general"][0]["q2"][ok], "plugin rollout q2", rtol=1e-7)
    print("damper_only 2^18 x 20 steps: table-driven %.2f ms, plug-in %.2f ms (%.1fx)" % (
        t["general"][1], t["plugin"][1], t["general"][1] / t["plugin"][1]))
    assert t["plugin"][1] < t["general"][1]


def test_cooperative_plugin_for_a_user_shape(lib):
    """kind="coop": the compile-time-size cooperative kernels for a shape the library was not built with (the rod
    system: fixed-length Distance + a kinematic slide)."""
    from trep_b200 import build
    name = "rod"
    d = G.desc(name)
    g = G.golden(name)
    assert lib.System(d, cooperative=True).kernel_name == "cooperative"        # run-time sizes before the plug-in
    path = os.path.join(os.path.dirname(build.LIB), "libtrepb_plugin_rod_coop.so")
    if shutil.which(build.NVCC) is not None:
        path = build.build_plugin(d, "rod_coop", kind="coop")
    elif not os.path.exists(path):
        pytest.skip("plug-in not prebuilt and no nvcc on this box")
    assert lib.load_plugin(path) == 1
    s = lib.System(d, cooperative=True)
    assert s.cooperative and s.kernel_name == "cooperative/rod_coop"
    out = s.linearize(g["case_q1"], g["case_p1"], g["case_u1"], g["case_k2"], t1=g["case_t1"], t2=g["case_t2"],
                      q2_guess=g["case_q2_guess"], lambda_guess=g["case_lambda_guess"], want_raw=True)
    assert np.all(out["status"] == 0)
    for k in ("q2", "p2", "lambda1", "A", "B"):
        G.assert_close(out[k], g["case_" + k], "coop plugin %s" % k)
    for k in G.RAW:
        G.assert_close(out[k], g["case_" + k], "coop plugin %s" % k)
    assert np.array_equal(out["iters"], g["case_iters"])


def test_cooperative_plugin_for_a_shape_with_springs(lib):
    """CtDims<..., 1>: compile-time sizes for everything but the spring / damper / wrench counts; built with
    ext=True, so the plug-in also carries the external-slab flavour of the linearize kernel (the default once
    loaded; TREPB_FLAG_COOP_ONE_WARP selects the plain one)."""
    from trep_b200 import build
    name = "spring_arms"
    d = G.desc(name)
    g = G.golden(name)
    assert lib.System(d, cooperative=True).kernel_name == "cooperative"
    path = os.path.join(os.path.dirname(build.LIB), "libtrepb_plugin_spring_arms_coop.so")
    if shutil.which(build.NVCC) is not None:
        path = build.build_plugin(d, "spring_arms_coop", kind="coop", ext=True)
    elif not os.path.exists(path):
        pytest.skip("plug-in not prebuilt and no nvcc on this box")
    assert lib.load_plugin(path) == 2
    for kw, kname in ((dict(), "cooperative/spring_arms_coop/ext"), (dict(coop_one_warp=True), "cooperative/spring_arms_coop")):
        s = lib.System(d, cooperative=True, **kw)
        assert s.cooperative and s.kernel_name == kname
        out = s.linearize(g["case_q1"], g["case_p1"], g["case_u1"], g["case_k2"], t1=g["case_t1"], t2=g["case_t2"],
                          q2_guess=g["case_q2_guess"], lambda_guess=g["case_lambda_guess"], want_raw=True)
        assert np.all(out["status"] == 0)
        for k in ("q2", "p2", "lambda1", "A", "B"):
            G.assert_close(out[k], g["case_" + k], "coop plugin %s %s" % (kname, k))
        for k in G.RAW:
            G.assert_close(out[k], g["case_" + k], "coop plugin %s %s" % (kname, k))
        assert np.array_equal(out["iters"], g["case_iters"])
    # the in-kernel stepping of the same flavour (solve-only layout with the run-time tail)
    dt, nsteps = float(g["roll_dt"]), int(g["roll_nsteps"])
    p0 = s.calc_p2(dt, g["roll_q0"], g["roll_q1"])
    st = s.step(g["roll_q1"], p0, dt, dt, nsteps=nsteps, u1=g["roll_u"][None], k2=g["roll_k2"][None], sample_every=1)
    assert st["status"][0] == 0
    G.assert_close(st["traj_q"][0], g["roll_q"][1:], "coop plugin traj q", rtol=1e-7)


def test_specialize_builds_and_loads_by_structure(lib):
    """lib.specialize: one call builds (cached by structure hash) and loads the plug-in; a description with the same
    structure and other numbers then runs on the same kernels."""
    from trep_b200 import build
    if shutil.which(build.NVCC) is None:
        pytest.skip("no nvcc on this box")
    name = "loop3d"
    d = G.desc(name)
    g = G.golden(name)
    before = lib.System(d)
    assert not before.specialized          # table-driven (the cooperative kernels built for this shape)
    assert lib.specialize(d, kind="thread") == 1
    assert lib.specialize(d, kind="thread") == 0
    s = lib.System(d)
    assert s.specialized and s.kernel_name.startswith("s") and s.kernel_name.endswith("_thread")
    kw = dict(t1=g["case_t1"], t2=g["case_t2"], q2_guess=g["case_q2_guess"], lambda_guess=g["case_lambda_guess"], want_raw=True)
    out = s.linearize(g["case_q1"], g["case_p1"], g["case_u1"], g["case_k2"], **kw)
    assert np.all(out["status"] == 0) and np.array_equal(out["iters"], g["case_iters"])
    for k in ("q2", "p2", "lambda1", "A", "B") + tuple(G.RAW):
        G.assert_close(out[k], g["case_" + k], "specialize %s" % k)
    rng = np.random.default_rng(3)
    B = 1 << 16
    idx = rng.integers(1, g["roll_q"].shape[0] - 1, B)
    q1 = g["roll_q"][idx] + rng.normal(0, 0.01, (B, d.nq)); p1 = g["roll_p"][idx] + rng.normal(0, 0.05, (B, d.nd))
    lam = g["roll_lambda"][idx - 1]
    t = {}
    for label, sysm in (("table-driven", before), ("specialised", s)):
        o = sysm.linearize(q1, p1, np.zeros((B, d.nu)), np.zeros((B, d.nk)), t1=0.0, t2=0.01, lambda_guess=lam)
        t[label] = (o, sysm.last_kernel_ms())
    ok = (t["table-driven"][0]["status"] == 0) & (t["specialised"][0]["status"] == 0)
    assert ok.mean() > 0.99
    G.assert_close(t["specialised"][0]["A"][ok], t["table-driven"][0]["A"][ok], "specialize A vs table-driven")
    print("loop3d 2^16 linearizations: %s %.2f ms, specialised %.2f ms" % (before.kernel_name, t["table-driven"][1], t["specialised"][1]))
