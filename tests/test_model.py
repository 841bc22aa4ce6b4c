"""Host-side model mirror: frame ordering, config ordering and flattening follow the reference's
synchronize step (trep/system.py:733-771, trep/frame.py:190-198)."""
import numpy as np

import golden_util as G
from trep_b200 import desc as D, model as M, systems


def test_frame_order_is_preorder_and_configs_dyn_then_kin():
    s = M.System()
    s.import_frames([
        M.rx("a"), [M.tz(-1, mass=1), M.ry("b", kinematic=True), [M.tx(2.0, name="tip", mass=2)]],
        M.ty("c")])
    d = s.describe()
    assert [c.name for c in s.configs] == ["a", "c", "b"]
    assert list(d.frame_parent) == [-1, 0, 1, 1, 3, 0]
    assert list(d.frame_config) == [-1, 0, -1, 2, -1, 1]
    assert d.nd == 2 and d.nk == 1
    assert d.ancestors(4) == [0, 2]
    assert d.mass_frames() == [2, 4]


def test_named_small_systems_sizes():
    want = {"pendulum1": (3, 1, 0, 0), "pendulum5": (11, 5, 0, 0), "damped_pendulum": (4, 1, 0, 0),
            "pend_on_cart1": (4, 2, 0, 1), "pend_on_cart2": (4, 2, 0, 2), "dual_pendulums": (6, 2, 0, 0),
            "pccd": (24, 7, 0, 0), "wrench_arm": (7, 3, 0, 5), "spline_pendulum": (5, 2, 0, 0)}
    for n, (nf, nd, nk, nu) in want.items():
        d = G.desc(n)
        assert (d.n_frames, d.nd, d.nk, d.nu) == (nf, nd, nk, nu), n


def test_puppet_description():
    d = systems.puppet_desc()
    assert (d.n_frames, d.nd, d.nk, d.nc, d.nu) == (86, 22, 18, 6, 0)
    assert len(d.mass_frames()) == 10
    assert d.nX == 80 and d.nU == 18
    # string-length configs drive no frame
    assert int(np.sum(d.config_frame() < 0)) == 6


def test_distance_constraint_with_kinematic_length():
    s = M.System()
    s.import_frames([M.tx("x", name="a", mass=1), M.ty("y", kinematic=True, name="b")])
    M.Distance(s, "a", "b", "len")
    M.PointToPoint2D(s, "xy", "a", "b")
    d = s.describe()
    assert d.nk == 2 and d.nc == 3
    assert list(d.con_kind) == [D.CON_DISTANCE, D.CON_POINT1D, D.CON_POINT1D]
    assert d.con_i[0, 2] == 2  # the length config comes last among the kinematic configs


def test_json_round_trip():
    for n in G.ALL:
        d = G.desc(n)
        assert D.SystemDesc.from_json(d.to_json()).equal(d)


def test_flatten_of_live_reference_systems_equals_the_mirror():
    """flatten_trep_system on the reference's own objects (oracle/_ref) gives the description the
    native model mirror gives - the drop-in route and the script route agree."""
    import os
    import sys
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(root, "oracle", "_ref", "trep")):
        pytest.skip("oracle/_ref not built")
    sys.path.insert(0, os.path.join(root, "oracle"))
    import ref_systems as R
    for n in G.ALL + G.EXTRA:
        d = M.flatten_trep_system(R.REF_BUILDERS[n](), name=n)
        assert d.equal(G.desc(n)), n


def test_unsupported_plugins_are_refused_not_approximated():
    import os
    import sys
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(root, "oracle", "_ref", "trep")):
        pytest.skip("oracle/_ref not built")
    sys.path.insert(0, os.path.join(root, "oracle"))
    import ref_systems as R
    trep = R.trep
    system = R.REF_BUILDERS["pendulum1"]()

    class PythonPotential(trep.Potential):      # a Python-defined plugin: no device implementation
        def V(self):
            return 0.0

    PythonPotential(system, "python-potential")
    with pytest.raises(TypeError):
        M.flatten_trep_system(system)
