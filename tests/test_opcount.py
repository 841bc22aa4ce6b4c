"""The operation-counting build of the thread-per-instance math (tests/opcount.cc) computes what the plain build
computes - so its counts are the counts of the algorithm the kernels run - and the counts recorded in
profiles/opcounts.json (tools/count_ops.py) are reproducible."""
import json
import os

import numpy as np
import pytest

import golden_util as G
import hostmath as H
import opcount as OC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["damped_pendulum", "pend_on_cart1", "dual_pendulums", "puppet"])
def test_counting_build_matches_plain_build(name):
    g = G.golden(name)
    d = G.desc(name)
    for c in range(min(3, g["case_q1"].shape[0])):
        args = (d, float(g["case_t1"][c]), float(g["case_t2"][c]), g["case_q1"][c], g["case_p1"][c], g["case_u1"][c],
                g["case_k2"][c])
        kw = dict(q2_guess=g["case_q2_guess"][c], lam_guess=g["case_lambda_guess"][c])
        a = OC.linearize(*args, **kw)
        b = H.linearize(*args, **kw)
        assert a["iters"] == b["iters"] == int(g["case_iters"][c])
        assert np.array_equal(a["A"], b["A"]) and np.array_equal(a["B"], b["B"])
        assert OC.flops(a["counts"]) > 0
        again = OC.linearize(*args, **kw)
        assert again["counts"] == a["counts"]


def test_recorded_counts_are_current():
    path = os.path.join(ROOT, "profiles", "opcounts.json")
    rec = json.load(open(path))
    d = G.desc("damped_pendulum")
    out = OC.step(d, 50, 0.0, 0.01, [0.7], [0.3])
    key = rec["damped_pendulum_step"]["check"]
    assert key["iters"] == out["iters"] and key["flops"] == OC.flops(out["counts"]) and key["sincos"] == out["counts"]["sincos"]
