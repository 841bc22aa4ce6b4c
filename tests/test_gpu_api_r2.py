"""GPU tests of the round-2 entry points, through the C ABI, against the reference run live (oracle/_ref):
trepb_calc_f_batch / trepb_discrete_fm2_batch (row a15), non-uniform time grids in the in-kernel loops
(`times` of trepb_step_args / trepb_project_args), trajectory-shaped linearize batches (`traj_len`), and the
ordering of launches that share a handle's scratch across streams."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from trep_b200 import lib as L
    assert L.device_count() > 0, "GPU tests need a CUDA device"
    return L


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_systems as R
    return R


def _flavours(lib, name):
    d = G.desc(name)
    out = [lib.System(d), lib.System(d, specialize=False, cooperative=False)]
    if lib.coop_dims(d) is not None:
        out.append(lib.System(d, specialize=False, cooperative=True))
    return out


@pytest.mark.parametrize("name", ["pend_on_cart2", "dual_pendulums", "fourbar", "rod", "puppet", "wrench_arm"])
def test_calc_f_and_discrete_fm2_match_reference(lib, ref, name):
    """_MidpointVI._calc_f (midpointvi.c:533-575) and discrete_fm2 (:2710-2727) at arbitrary - not solved -
    (q1, q2, p1, u1, lambda1): the residual a caller inspects, and the discrete forcing."""
    rng = np.random.default_rng(17)
    system, mvi = ref.make_mvi(name)
    nq, nd, nu, nc = mvi.nq, mvi.nd, mvi.nu, mvi.nc
    B = 9
    g = G.golden(name)
    base = g["roll_q"][rng.integers(0, g["roll_q"].shape[0], B)] if "roll_q" in g and g["roll_q"].shape[1] == nq \
        else rng.uniform(-1, 1, (B, nq))
    q1 = base + rng.normal(0, 0.05, (B, nq)); q2 = q1 + rng.normal(0, 0.02, (B, nq))
    p1 = rng.normal(0, 1, (B, nd)); u1 = rng.uniform(-1, 1, (B, nu)); lam = rng.normal(0, 1, (B, nc))
    t1, t2 = 0.3, 0.3125
    want_f = np.zeros((B, nd + nc)); want_fm2 = np.zeros((B, nd))
    for b in range(B):
        mvi.t1, mvi.t2 = t1, t2
        mvi.q1, mvi.q2, mvi.p1 = q1[b], q2[b], p1[b]
        if nu: mvi.u1 = u1[b]
        if nc: mvi.lambda1 = lam[b]
        mvi._calc_f()
        want_f[b] = np.array(mvi._f)
        want_fm2[b] = mvi.discrete_fm2()
    for s in _flavours(lib, name):
        f = s.calc_f(t1, t2, q1, q2, p1, u1, lam)
        fm2 = s.discrete_fm2(t1, t2, q1, q2, u1)
        G.assert_close(f, want_f, "%s[%s] f" % (name, s.kernel_name))
        G.assert_close(fm2, want_fm2, "%s[%s] fm2" % (name, s.kernel_name))


def test_calc_f_of_a_solved_step_is_below_the_tolerance(lib):
    s = lib.System(G.desc("pend_on_cart1"))
    rng = np.random.default_rng(2)
    q1 = rng.uniform(-1, 1, (64, 2)); p1 = rng.normal(0, 1, (64, 2)); u1 = rng.uniform(-1, 1, (64, 1))
    out = s.step(q1, p1, 0.0, 0.01, u1=u1[:, None, :])
    f = s.calc_f(0.0, 0.01, q1, out["q2"], p1, u1)
    assert np.all(np.linalg.norm(f, axis=1) <= 1e-10)


@pytest.mark.parametrize("name", ["pend_on_cart1", "rod", "puppet"])
def test_nonuniform_time_grid_rollout_matches_reference(lib, ref, name):
    """The reference steps on arbitrary self._time[k] (dsystem.py:229-250); the in-kernel loop takes the grid."""
    rng = np.random.default_rng(23)
    system, mvi = ref.make_mvi(name)
    nq, nd, nu, nk = mvi.nq, mvi.nd, mvi.nu, mvi.nk
    g = G.golden(name)
    K = 25
    times = np.concatenate([[0.0], np.cumsum(rng.uniform(0.004, 0.016, K))])
    if name == "pend_on_cart1":
        q0 = np.array([0.1, 0.4]); u = rng.uniform(-1, 1, (K, nu)); k = np.zeros((K, 0))
    else:
        q0 = g["roll_q0"]
        u = np.zeros((K, 0))
        k = g["roll_k2"][:K] if name == "puppet" else q0[nd:] + 0.05 * np.sin(40 * times[1:, None])
    mvi.initialize_from_configs(0.0, q0, times[0] + 1e-2, q0)
    mvi.t2 = times[0]
    p0 = mvi.p2
    want_q, want_p, its = [], [], 0
    for s_ in range(K):
        its += mvi.step(times[s_ + 1], tuple(u[s_]), tuple(k[s_]))
        want_q.append(mvi.q2); want_p.append(mvi.p2)
    for s in _flavours(lib, name):
        out = s.step(q0, p0, 0.0, 0.01, nsteps=K, u1=u[None] if nu else None, k2=k[None] if nk else None,
                     sample_every=1, times=times)
        assert out["status"][0] == 0
        G.assert_close(out["traj_q"][0], np.array(want_q), "%s[%s] q" % (name, s.kernel_name), rtol=1e-8)
        G.assert_close(out["traj_p"][0], np.array(want_p), "%s[%s] p" % (name, s.kernel_name), rtol=1e-8)
        assert abs(int(out["iters"][0]) - its) <= 1


def test_project_on_a_nonuniform_grid_matches_reference(lib, ref):
    """DSystem.project (dsystem.py:426-457) on a non-uniform time base, reference classes vs the batched mirror."""
    from trep_b200 import discopt, midpointvi
    rng = np.random.default_rng(5)
    K = 30
    t = np.concatenate([[0.0], np.cumsum(rng.uniform(0.005, 0.015, K))])
    system, mvi = ref.make_mvi("pend_on_cart1")
    rds = ref.discopt.DSystem(mvi, t)
    bU = 0.5 * np.sin(3 * t[:-1])[:, None]
    bX = np.zeros((K + 1, 4)); bX[:, 1] = 0.2 * np.cos(2 * t); bX[:, 0] = 0.1 * t
    Kp = rng.normal(0, 0.3, (K, 1, 4))
    want = rds.project(bX, bU, Kp)
    d = discopt.DSystem(midpointvi.MidpointVI(G.desc("pend_on_cart1")), t)
    got = d.project(bX, bU, Kp)
    G.assert_close(got.X, want[0], "projected X", rtol=1e-9)
    G.assert_close(got.U, want[1], "projected U", rtol=1e-9)
    # open loop with the same inputs reproduces the states (to the Newton tolerance: no hints here)
    X2 = d.simulate(bX[:1], got.U[None])
    G.assert_close(X2[0], got.X, "simulate", rtol=1e-7)


@pytest.mark.parametrize("name", ["pend_on_cart1", "rod"])
def test_trajectory_shaped_linearize_equals_per_step_instances(lib, name):
    """traj_len: R trajectories of L state rows -> R (L-1) linearizations with X[k+1] as the Newton start, outputs
    packed [R][L-1]; bit-identical to the flat batch of the same (state, input, hint) triples."""
    s = lib.System(G.desc(name))
    rng = np.random.default_rng(3)
    R, L = 5, 12
    nq, nd, nu, nk = s.nq, s.nd, s.nu, s.nk
    g = G.golden(name)
    q0 = np.tile(g["roll_q0"] if name == "rod" else np.array([0.0, 0.3]), (R, 1)) + rng.normal(0, 1e-3, (R, nq))
    if name == "rod":
        q0 = np.tile(g["roll_q0"], (R, 1))
    u = rng.uniform(-1, 1, (R, L - 1, nu))
    k = np.tile(g["roll_k2"][:L - 1], (R, 1, 1)) if nk else None
    p0 = s.calc_p2(0.01, q0, q0)
    roll = s.step(q0, p0, 0.0, 0.01, nsteps=L - 1, u1=u if nu else None, k2=k, sample_every=1)
    assert np.all(roll["status"] == 0)
    Q = np.concatenate([q0[:, None], roll["traj_q"]], axis=1)       # [R][L][nq]
    P = np.concatenate([p0[:, None], roll["traj_p"]], axis=1)
    n = R * (L - 1)
    up = lambda a, dt=np.float64: lib.DeviceBuffer(0, a.shape, dt).upload(np.ascontiguousarray(a, dtype=dt))

    class Off:
        def __init__(self, buf, off_bytes): self.p = buf.data_ptr() + off_bytes
        def data_ptr(self): return self.p
    dQ, dP, dH = up(Q), up(P), up(Q[:, :, :nd])
    dU = up(u.reshape(n, nu)) if nu else None
    dK = up(k.reshape(n, nk)) if nk else None
    dA, dB = lib.DeviceBuffer(0, (n, s.nX, s.nX)), lib.DeviceBuffer(0, (n, s.nX, max(s.nU, 1)))
    dst, dit = lib.DeviceBuffer(0, (n,), np.int32), lib.DeviceBuffer(0, (n,), np.int32)
    # X[k+1] as the Newton start of step k: the hint array shifted by one row
    s.linearize_raw(True, n, dQ, dP, dU, dK, dst, t1_scalar=0.0, dt_scalar=0.01, q2_guess=Off(dH, nd * 8), iters=dit,
                    A=dA, B=dB if s.nU else None, traj_len=L)
    lib.synchronize(0)
    flat = s.linearize(Q[:, :-1].reshape(n, nq), P[:, :-1].reshape(n, nd), u.reshape(n, nu) if nu else None,
                       k.reshape(n, nk) if nk else None, t1=0.0, dt=0.01, q2_guess=Q[:, 1:, :nd].reshape(n, nd))
    assert np.all(dst.download() == 0) and np.array_equal(dit.download(), flat["iters"])
    assert np.array_equal(dA.download(), flat["A"])
    if s.nU:
        assert np.array_equal(dB.download()[:, :, :s.nU].reshape(flat["B"].shape), flat["B"])
    assert flat["iters"].max() <= 1          # the hint is the solution: nothing (or one polishing step) left to do


def test_one_handle_on_two_streams_keeps_its_scratch_consistent(lib):
    """launches of the table-driven kernels share the handle's workspace slab; issued on two streams
    they must still give the results of the serial order (the second launch waits for the first on the device)."""
    raw = lib.raw()
    cudart = C.CDLL("libcudart.so.12")
    s = lib.System(G.desc("pccd"), specialize=False, cooperative=False)
    g = G.golden("pccd")
    rng = np.random.default_rng(1)
    B = 20000
    idx = rng.integers(1, g["roll_q"].shape[0] - 1, (2, B))
    st = [C.c_void_p(), C.c_void_p()]
    for x in st:
        assert cudart.cudaStreamCreateWithFlags(C.byref(x), 1) == 0      # non-blocking streams
    up = lambda a, dt=np.float64: lib.DeviceBuffer(0, a.shape, dt).upload(np.ascontiguousarray(a, dtype=dt))
    bufs, outs = [], []
    for j in range(2):
        q1 = g["roll_q"][idx[j]] + rng.normal(0, 0.01, (B, s.nq)); p1 = g["roll_p"][idx[j]]
        lam = g["roll_lambda"][idx[j] - 1]
        bufs.append((q1, p1, lam, up(q1), up(p1), up(lam)))
        outs.append((lib.DeviceBuffer(0, (B, s.nX, s.nX)), lib.DeviceBuffer(0, (B,), np.int32)))
    for rep in range(3):
        for j in range(2):
            q1, p1, lam, dq, dp, dl = bufs[j]
            s.linearize_raw(True, B, dq, dp, None, None, outs[j][1], t1_scalar=0.0, dt_scalar=0.01, lambda_guess=dl,
                            A=outs[j][0], stream=st[j])
    lib.synchronize(0)
    for j in range(2):
        q1, p1, lam, *_ = bufs[j]
        want = s.linearize(q1, p1, t1=0.0, dt=0.01, lambda_guess=lam)
        assert np.array_equal(outs[j][1].download(), want["status"])
        assert np.array_equal(outs[j][0].download(), want["A"])
    for x in st:
        cudart.cudaStreamDestroy(x)


def test_specialised_kernels_take_parameters_at_run_time(lib, ref):
    """Structure is compiled in, numbers are kernel arguments: damped pendulums with other masses, lengths,
    pivots, gravity and damping run the kernel built for examples/damped-pendulum.py (not the table-driven
    fallback) and match the reference built with the same numbers."""
    from trep_b200 import model as M
    trep = ref.trep
    rng = np.random.default_rng(77)
    for m, l, y0, grav, c in ((1.0, 3.0, 3.0, -9.8, 1.2), (2.0, 5.0, 1.0, -9.8, 1.2), (0.3, 0.7, -2.0, -1.62, 0.05),
                              (7.5, 2.2, 4.0, -24.8, 3.3)):
        s = M.System(name="damped_pendulum")
        s.import_frames([M.ty(y0), M.rx("theta"), [M.tz(-l, mass=m)]])
        M.Gravity(s, (0, 0, grav))
        M.Damping(s, c)
        h = lib.System(s.describe())
        assert h.specialized and h.kernel_name == "damped_pendulum"
        rs = trep.System()
        rs.import_frames([trep.ty(y0), trep.rx("theta"), [trep.tz(-l, mass=m)]])
        trep.potentials.Gravity(rs, (0, 0, grav))
        trep.forces.Damping(rs, c)
        mvi = trep.MidpointVI(rs, num_threads=1)
        B = 64
        q1 = rng.uniform(-3, 3, (B, 1)); p1 = rng.normal(0, 2 * m * l * l, (B, 1))
        want = ref.run_cases(mvi, 0.0, 0.01, q1, p1, np.zeros((B, 0)), np.zeros((B, 0)))
        got = h.linearize(q1, p1, t1=0.0, t2=0.01)
        assert np.array_equal(got["status"], want["status"]) and np.array_equal(got["iters"], want["iters"])
        for k in ("q2", "p2", "A"):
            G.assert_close(got[k], want[k], "m=%g l=%g %s" % (m, l, k))
        # second derivatives through the specialised per-pair kernel
        mvi.initialize_from_state(0.0, q1[0], p1[0]); mvi.step(0.01); mvi._calc_deriv2()
        d2 = h.deriv2(q1[:1], p1[:1], t1=0.0, t2=0.01)
        G.assert_close(d2["q2_dq1dq1"][0], np.array(mvi._q2_dq1dq1), "m=%g l=%g q2_dq1dq1" % (m, l))
        G.assert_close(d2["p2_dq1dp1"][0], np.array(mvi._p2_dq1dp1), "m=%g l=%g p2_dq1dp1" % (m, l))
    # a structural change (a rotational inertia on the bob) is another kernel: table-driven here
    s = M.System()
    s.import_frames([M.ty(3), M.rx("theta"), [M.tz(-3, mass=(1.0, 0.2, 0.0, 0.0))]])
    M.Gravity(s, (0, 0, -9.8)); M.Damping(s, 1.2)
    assert not lib.System(s.describe()).specialized


def test_large_host_batches_are_pipelined_in_chunks_with_identical_results(lib):
    """trepb_step_batch / trepb_linearize_batch cut a large host batch into chunks on two streams (copies of one
    chunk overlap the kernel of another): same bits as the device-resident call on the whole batch, ragged
    chunk boundaries included."""
    from trep_b200 import systems
    rng = np.random.default_rng(5)
    d = systems.named_desc("pend_on_cart1")
    s = lib.System(d)
    B = 4 * 65536 + 1237
    q1 = rng.uniform(-2, 2, (B, d.nq)); p1 = rng.normal(0, 1, (B, d.nd)); u1 = rng.uniform(-1, 1, (B, 3, d.nu))
    host = s.step(q1, p1, 0.0, 0.01, nsteps=3, u1=u1, sample_every=1)
    up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(np.ascontiguousarray(a))
    dq, dp, du = up(q1), up(p1), up(u1)
    q2 = lib.DeviceBuffer(0, (B, d.nq)); p2 = lib.DeviceBuffer(0, (B, d.nd))
    it = lib.DeviceBuffer(0, (B,), np.int32); st = lib.DeviceBuffer(0, (B,), np.int32)
    tq = lib.DeviceBuffer(0, (B, 3, d.nq)); tp = lib.DeviceBuffer(0, (B, 3, d.nd))
    s.step_raw(True, B, 3, 0.0, 0.01, dq, dp, du, None, None, None, q2, p2, None, it, st, sample_every=1, traj_q=tq, traj_p=tp)
    lib.synchronize(0)
    assert np.array_equal(host["q2"], q2.download()) and np.array_equal(host["p2"], p2.download())
    assert np.array_equal(host["iters"], it.download()) and np.array_equal(host["status"], st.download())
    assert np.array_equal(host["traj_q"], tq.download()) and np.array_equal(host["traj_p"], tp.download())
    lin = s.linearize(q1, p1, u1[:, 0], None, t1=np.zeros(B), t2=np.full(B, 0.01))
    A = lib.DeviceBuffer(0, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(0, (B, d.nX, d.nU))
    s.linearize_raw(True, B, dq, dp, up(u1[:, 0]), None, st, t1_scalar=0.0, dt_scalar=0.01, q2=q2, p2=p2, iters=it, A=A, B=Bm)
    lib.synchronize(0)
    assert np.array_equal(lin["A"], A.download()) and np.array_equal(lin["B"], Bm.download())
    assert np.array_equal(lin["q2"], q2.download()) and np.array_equal(lin["status"], st.download())
