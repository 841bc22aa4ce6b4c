"""Helpers shared by the parity tests: golden fixture access and the parity metric.

Golden vectors are outputs of the reference itself (oracle/_ref, built from /root/reference by
oracle/build_ref.py) recorded by oracle/gen_golden.py; see tests/golden/."""
import os

import numpy as np

from trep_b200 import systems

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SMALL = ["tase_pendulum", "pendulum1", "pendulum5", "damped_pendulum", "pend_on_cart1",
         "pend_on_cart2", "dual_pendulums"]
ALL = SMALL + ["puppet"]
# plugin kinds beyond BASELINE.json's configs: PointOnPlane constraints, Body/Hybrid/Spatial wrenches,
# NonlinearConfigSpring over a spline (fixtures from oracle/gen_golden_f4.py)
EXTRA = ["pccd", "wrench_arm", "spline_pendulum"]
# ... with second-derivative goldens (the reference's third derivative of the spline spring has the
# wrong sign, potentials/nonlinear_config_spring.c:53, so none is recorded for it)
EXTRA_D2 = ["pccd", "wrench_arm"]

RAW = ["q2_dq1", "q2_dp1", "q2_du1", "q2_dk2", "p2_dq1", "p2_dp1", "p2_du1", "p2_dk2",
       "l1_dq1", "l1_dp1", "l1_du1", "l1_dk2"]

# north_star: "match the reference _trep on identical inputs to within 1e-10 relative in fp64"
RTOL = 1e-10


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def desc(name):
    return systems.named_desc(name)


def relerr(a, b):
    """max |a-b| relative to the magnitude of the reference array (floor 1)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(1.0, float(np.max(np.abs(b))))
    return float(np.max(np.abs(a - b))) / scale


def assert_close(a, b, what, rtol=RTOL):
    e = relerr(a, b)
    assert e <= rtol, "%s: relative error %.3e > %.1e" % (what, e, rtol)
