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
