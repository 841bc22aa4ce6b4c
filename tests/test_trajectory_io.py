"""Trajectory file round trips, and interchange with the reference's own save_trajectory /
load_trajectory (trep/system.py:1209-1304) when oracle/_ref is built."""
import os
import sys

import numpy as np
import pytest

import golden_util as G
from trep_b200 import trajectory as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _traj(d, K, rng, batch=None):
    lead = () if batch is None else (batch,)
    return dict(t=0.01 * np.arange(K), Q=rng.normal(size=lead + (K, d.nq)), p=rng.normal(size=lead + (K, d.nd)),
                v=rng.normal(size=lead + (K, d.nk)), u=rng.normal(size=lead + (K - 1, d.nu)),
                rho=rng.normal(size=lead + (K - 1, d.nk)))


@pytest.mark.parametrize("name", ["pend_on_cart2", "puppet"])
@pytest.mark.parametrize("batch", [None, 3])
def test_round_trip(tmp_path, name, batch):
    d = G.desc(name)
    x = _traj(d, 7, np.random.default_rng(0), batch)
    f = str(tmp_path / "traj.mat")
    T.save_trajectory(f, d, x["t"], x["Q"], x["p"], x["v"] if d.nk else None, x["u"] if d.nu else None,
                      x["rho"] if d.nk else None)
    t, Q, p, v, u, rho = T.load_trajectory(f, d)
    assert np.array_equal(t, x["t"]) and np.array_equal(Q, x["Q"]) and np.array_equal(p, x["p"])
    if d.nk:
        assert np.array_equal(v, x["v"]) and np.array_equal(rho, x["rho"])
    else:
        assert v is None and rho is None
    if d.nu:
        assert np.array_equal(u, x["u"])
    # without a system: names + arrays as stored
    t2, (qi, Qraw), (pi, _), _, (ui, _), _ = T.load_trajectory(f)
    assert qi == list(d.config_names) and pi == list(d.config_names[:d.nd]) and ui == list(d.input_names)
    assert np.array_equal(Qraw, x["Q"])


def test_columns_are_matched_by_name(tmp_path):
    """A file written for one system loads into another that shares some names: shared columns are
    copied, the others stay zero (system.py:1262-1302)."""
    a, b = G.desc("pend_on_cart1"), G.desc("pend_on_cart2")
    x = _traj(b, 5, np.random.default_rng(1))
    f = str(tmp_path / "traj.mat")
    T.save_trajectory(f, b, x["t"], x["Q"], x["p"], None, x["u"], None)
    t, Q, p, v, u, rho = T.load_trajectory(f, a)
    assert np.array_equal(Q, x["Q"]) and u.shape == (4, 1) and np.array_equal(u[:, 0], x["u"][:, 0])


def test_interchange_with_the_reference(tmp_path):
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "trep")):
        pytest.skip("oracle/_ref not built")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_systems as R
    trep = R.trep
    system = R.REF_BUILDERS["puppet"]()
    d = G.desc("puppet")
    x = _traj(d, 6, np.random.default_rng(2))
    # written here, read by the reference
    f = str(tmp_path / "ours.mat")
    T.save_trajectory(f, d, x["t"], x["Q"], x["p"], x["v"], None, x["rho"])
    t, Q, p, v, u, rho = trep.load_trajectory(f, system)
    assert np.array_equal(Q, x["Q"]) and np.array_equal(p, x["p"]) and np.array_equal(v, x["v"])
    assert np.array_equal(rho, x["rho"]) and np.allclose(t, x["t"])
    # written by the reference, read here (with the description and with the live system)
    f = str(tmp_path / "theirs.mat")
    trep.save_trajectory(f, system, x["t"], x["Q"], x["p"], x["v"], None, x["rho"])
    for target in (d, system):
        t, Q, p, v, u, rho = T.load_trajectory(f, target)
        assert np.array_equal(Q, x["Q"]) and np.array_equal(p, x["p"]) and np.array_equal(v, x["v"])
        assert np.array_equal(rho, x["rho"]) and u is None


class _FakeVarint:
    """The trajectory-packing half of the DSystem mirror needs only sizes and names."""
    def __init__(self, d):
        self.desc = d
        self.nq, self.nd, self.nk, self.nu = d.nq, d.nd, d.nk, d.nu


def _mirror(name, K):
    from trep_b200.discopt import DSystem
    return DSystem(_FakeVarint(G.desc(name)), 0.01 * np.arange(K + 1))


def test_dsystem_trajectory_helpers_match_the_reference(tmp_path):
    """build_trajectory / split_trajectory / save_state_trajectory / load_state_trajectory /
    convert_trajectory / dproject of the mirror against the reference's DSystem
    (trep/discopt/dsystem.py:140-226, 388-402, 460-471, 497-534)."""
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "trep")):
        pytest.skip("oracle/_ref not built")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_systems as R
    trep, discopt = R.trep, R.discopt
    K = 9
    t = 0.01 * np.arange(K + 1)
    rng = np.random.default_rng(3)
    refs, mirrors = {}, {}
    for name in ("puppet", "pend_on_cart1", "pend_on_cart2"):
        system, mvi = R.make_mvi(name)
        refs[name] = discopt.DSystem(mvi, t)
        mirrors[name] = _mirror(name, K)
    ref, mir = refs["puppet"], mirrors["puppet"]
    d = G.desc("puppet")
    x = _traj(d, K + 1, rng)
    Xr, Ur = ref.build_trajectory(x["Q"], x["p"], x["v"], None, x["rho"])
    Xm, Um = mir.build_trajectory(x["Q"], x["p"], x["v"], None, x["rho"])
    assert np.array_equal(Xr, Xm) and np.array_equal(Ur, Um)
    for a, b in zip(ref.split_trajectory(Xr, Ur), mir.split_trajectory(Xm, Um)):
        assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        mir.build_trajectory(Q=x["Q"][:-1])
    # files: written by the mirror, read by the reference, and back
    f = str(tmp_path / "state.mat")
    mir.save_state_trajectory(f, Xm, Um)
    X2, U2 = ref.load_state_trajectory(f)
    assert np.array_equal(X2, Xr) and np.array_equal(U2, Ur)
    ref.save_state_trajectory(f, Xr, Ur)
    X3, U3 = mir.load_state_trajectory(f)
    assert np.array_equal(X3, Xr) and np.array_equal(U3, Ur) and np.allclose(mir.time, t)
    # convert_trajectory between systems that share some names
    a, b = "pend_on_cart2", "pend_on_cart1"
    da = G.desc(a)
    xa = _traj(da, K + 1, rng)
    Xa, Ua = refs[a].build_trajectory(xa["Q"], xa["p"], None, xa["u"], None)
    # (the reference's own convert_trajectory indexes with dict views, Python-2 only, so the expectation
    # is written out: both systems share x, theta and the input x-force; theta-force has no counterpart)
    got = mirrors[b].convert_trajectory(mirrors[a], Xa, Ua)
    assert np.array_equal(got.X, Xa) and got.U.shape == (K, 1) and np.array_equal(got.U[:, 0], Ua[:, 0])
    back = mirrors[a].convert_trajectory(mirrors[b], got.X, got.U)
    assert np.array_equal(back.X, Xa) and np.array_equal(back.U[:, 0], Ua[:, 0]) and not back.U[:, 1].any()
    # dproject: one trajectory against the reference, and a batch against the single-trajectory results
    nX, nU = ref.nX, ref.nU
    A = rng.normal(0, 0.2, (3, K, nX, nX)); B = rng.normal(0, 0.2, (3, K, nX, nU))
    Kf = rng.normal(0, 0.1, (3, K, nU, nX)); bdX = rng.normal(size=(3, K + 1, nX)); bdU = rng.normal(size=(3, K, nU))
    got = mir.dproject(A, B, bdX, bdU, Kf)
    for i in range(3):
        want = ref.dproject(A[i], B[i], bdX[i], bdU[i], Kf[i])
        assert np.allclose(got.dX[i], want.dX, rtol=1e-12, atol=1e-12) and np.allclose(got.dU[i], want.dU, rtol=1e-12, atol=1e-12)
