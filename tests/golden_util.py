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
# constraint / force kinds no BASELINE config exercises (fixtures from oracle/gen_golden_r2.py): PointToPoint2D
# (fourbar), PointToPoint3D (loop3d), fixed-length Distance + a kinematic config (rod), LinearDamper alone
# (damper_only: the reference's _calc_deriv2 runs there)
PARITY = ["fourbar", "loop3d", "rod", "damper_only"]
PARITY_CONSTRAINED = ["fourbar", "loop3d", "rod"]
# LinearSprings between two arms, a kinematic slide, an input and a distance constraint: the spring terms of the
# cooperative kernels (no second-derivative goldens: the reference has no C V_dqdqdq for a LinearSpring)
PARITY_SPRING = ["spring_arms"]

RAW = ["q2_dq1", "q2_dp1", "q2_du1", "q2_dk2", "p2_dq1", "p2_dp1", "p2_du1", "p2_dk2",
       "l1_dq1", "l1_dp1", "l1_du1", "l1_dk2"]

# north_star: "match the reference _trep on identical inputs to within 1e-10 relative in fp64"
RTOL = 1e-10


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def desc(name):
    return systems.named_desc(name)


# element-wise bound: |a - b| <= RTOL |b| + ATOL_FRAC max|b|.  The second term covers entries that are small
# only through cancellation (their absolute error is set by the large terms they were formed from); it is tied to
# the array's OWN scale, not to 1, so an array whose largest entry is 2.5e-5 is held to 2.5e-16, not 1e-10.
# (1e-11 rather than something smaller: e.g. pend-on-cart's d p2_x / d x1 is a structural zero that both the
# reference (1.3e-14) and the kernels (2.3e-13) return as the rounding residue of cancelling m/dt-sized
# terms ~1e3, in an array whose largest entry is 0.085.)
ATOL_FRAC = 1e-11


def relerr(a, b, atol_frac=ATOL_FRAC):
    """Largest element-wise |a-b| / (|b| + (atol_frac / RTOL) max|b|): the value assert_close compares with RTOL."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = float(np.max(np.abs(b)))
    den = np.abs(b) + (atol_frac / RTOL) * scale
    diff = np.abs(a - b)
    if scale == 0.0:
        return 0.0 if float(np.max(diff)) == 0.0 else float("inf")
    with np.errstate(invalid="ignore"):
        r = diff / den
    r = np.where(np.isnan(diff), np.inf, r)
    return float(np.max(r))


def assert_close(a, b, what, rtol=RTOL, atol_frac=None, scale=None):
    """|a - b| <= rtol |b| + atol_frac max|b| element by element (atol_frac scales with rtol by default).
    `scale` replaces max|b| (see assert_d2_close)."""
    if atol_frac is None:
        atol_frac = ATOL_FRAC * (rtol / RTOL)
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size == 0:
        return
    if scale is None:
        scale = float(np.max(np.abs(b)))
    bound = rtol * np.abs(b) + atol_frac * scale
    diff = np.abs(a - b)
    bad = ~(diff <= bound)
    if np.any(bad):
        i = np.unravel_index(int(np.argmax(np.where(bad, diff - bound, -np.inf))), diff.shape)
        raise AssertionError("%s: |a-b| = %.3e at %s exceeds %.1e |b| + %.1e max|b| (b = %.3e, max|b| = %.3e)"
                             % (what, diff[i], i, rtol, atol_frac, b[i], scale))


D2_KINDS = ["dq1dq1", "dq1dp1", "dq1du1", "dq1dk2", "dp1dp1", "dp1du1", "dp1dk2", "du1du1", "du1dk2", "dk2dk2"]


def assert_d2_close(out, gold, what, names=None, rtol=RTOL, gold_prefix="case_", index=None):
    """The 30 second-derivative tensors: element-wise |a-b| <= rtol |b| + ATOL_FRAC S with S the largest entry of
    the tensor's output family (all q2_d.d., all p2_d.d. or all l1_d.d. of the compared cases).  A single tensor
    can be structurally zero in the reference (it skips the terms: exact 0.0) while the kernels return the
    rounding residue of the cancelled terms - e.g. rod's q2_dk1dp1 = 2^-42 - so the floor is the family's scale."""
    for fam in ("q2", "p2", "l1"):
        fam_names = [fam + "_" + k for k in D2_KINDS]
        pick = (lambda n: gold[gold_prefix + n] if index is None else gold[gold_prefix + n][index])
        S = max([float(np.max(np.abs(pick(n)))) for n in fam_names if pick(n).size] or [0.0])
        for n in fam_names:
            if names is not None and n not in names:
                continue
            assert_close(out[n], pick(n), "%s %s" % (what, n), rtol=rtol, scale=S)
