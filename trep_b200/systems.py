"""The named systems of BASELINE.json's configs, built with the model mirror.

Each builder cites the reference script it restates.  The marionette (config 5) is loaded
from a flattened description (``trep_b200/data/puppet.json``) that was produced from the
reference's own ``trep/puppets/puppets.py`` by ``oracle/gen_golden.py`` — re-typing the
86-frame skeleton here would only be a copy of data.
"""
from __future__ import annotations

import math
import os

from . import model as M
from .desc import SystemDesc

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def pendulum(links=1) -> M.System:
    """N-link pendulum: RX joint + TZ(-1) point mass per link, default gravity.
    (examples/pendulum.py:36-71)"""
    s = M.System(name="pendulum%d" % links)
    M.Gravity(s, name="Gravity")
    frame = s.world_frame
    for link in range(links):
        frame = M.Frame(frame, M.D.RX, "link-%d" % link, "link-%d" % link)
        frame = M.Frame(frame, M.D.TZ, -1)
        frame.set_mass(1.0)
    s.get_config("link-0").q = math.pi / 4.0
    return s


def damped_pendulum() -> M.System:
    """(examples/damped-pendulum.py:13-20)"""
    s = M.System(name="damped_pendulum")
    s.import_frames([M.ty(3), M.rx("theta"), [M.tz(-3, mass=1)]])
    M.Gravity(s, (0, 0, -9.8))
    M.Damping(s, 1.2)
    return s


def pend_on_cart(torque_force=False) -> M.System:
    """(examples/pend-on-cart-optimization.py:48-64)"""
    s = M.System(name="pend_on_cart%d" % (2 if torque_force else 1))
    s.import_frames([
        M.tx("x", name="Cart", mass=10.0), [
            M.rz("theta", name="PendulumBase"), [
                M.ty(-1.0, name="Pendulum", mass=1.0)]]])
    M.Gravity(s, (0, -9.8, 0))
    M.Damping(s, 0.01)
    M.ConfigForce(s, "x", "x-force")
    if torque_force:
        M.ConfigForce(s, "theta", "theta-force")
    return s


def dual_pendulums() -> M.System:
    """(examples/dual_pendulums.py:28-43)"""
    s = M.System(name="dual_pendulums")
    s.import_frames([
        M.rx("theta1"), [M.tz(2, mass=1, name="pend1")],
        M.ty(1), [M.rx("theta2"), [M.tz(2, mass=1, name="pend2")]]])
    M.LinearSpring(s, "pend1", "pend2", k=20, x0=1)
    M.LinearDamper(s, "pend1", "pend2", c=1)
    M.Gravity(s, name="Gravity")
    s.q = [3, -3]
    return s


def tase_pendulum() -> M.System:
    """Known-answer pendulum of examples/papers/tase2012/pend-single-step.py:8-27."""
    s = M.System(name="tase_pendulum")
    s.import_frames([M.rz("theta_1", name="PendAngle"), [M.ty(-1.0, name="PendMass", mass=1.0)]])
    M.Gravity(s, (0, -9.8, 0))
    M.ConfigForce(s, "theta_1", "tau")
    return s


def pccd() -> M.System:
    """Planar closed-chain device: three open chains closed by four PointOnPlane constraints
    (examples/pccd.py:20-66).  nd 7, nc 4."""
    s = M.System(name="pccd")
    s.import_frames([
        M.rx("J", name="J"), [
            M.tz(-0.5, name="I", mass=1),
            M.tz(-1), [
                M.rx("H", name="H"), [
                    M.tz(-1, name="G", mass=1),
                    M.tz(-2, name="O2")]]],
        M.ty(1.5), [
            M.rx("K", name="K"), [
                M.tz(-1, name="L", mass=1),
                M.tz(-2), [
                    M.rx("M", name="M"), [
                        M.tz(-0.5, name="N", mass=1),
                        M.tz(-1.0, name="O")]]]],
        M.ty(-1.5), [
            M.rx("A", name="A"), [
                M.tz(-1, name="B", mass=1),
                M.tz(-2), [
                    M.rx("C", name="C"), [
                        M.tz(-0.375, name="D", mass=1),
                        M.tz(-0.75), [
                            M.rx("E", name="E"), [
                                M.tz(-0.5, name="F", mass=1),
                                M.tz(-1.0, name="G2")]]]]]]])
    M.Gravity(s, (0, 0, -9.8))
    M.Damping(s, 0.1)
    M.PointOnPlane(s, "O", (0, 1, 0), "O2")
    M.PointOnPlane(s, "O", (0, 0, 1), "O2")
    M.PointOnPlane(s, "G", (0, 1, 0), "G2")
    M.PointOnPlane(s, "G", (0, 0, 1), "G2")
    return s


def wrench_arm() -> M.System:
    """Spatial three-joint arm (two revolute joints about different axes and a prismatic one, an
    offset constant frame) loaded by one wrench of each kind with mixed constant / input components
    (trep/forces/bodywrench.py, hybridwrench.py, spatialwrench.py; cf. the HybridWrench loads of
    examples/extensor-tendon-model.py:49-56).  nd 3, nu 5."""
    s = M.System(name="wrench_arm")
    s.import_frames([
        M.rz("yaw", name="base"), [
            M.tz(0.4, name="shoulder", mass=2.0), [
                M.ry("pitch", name="upper"), [
                    M.tx(0.7, name="elbow", mass=1.5), [
                        M.tx("reach", name="slider"), [
                            M.const_txyz((0.1, -0.2, 0.3), name="tool", mass=0.5)]]]]]])
    s.get_frame("shoulder").set_mass(2.0, 0.1, 0.2, 0.3)
    s.get_frame("tool").set_mass(0.5, 0.02, 0.03, 0.01)
    M.Gravity(s, (0, 0, -9.8))
    M.Damping(s, 0.05)
    M.BodyWrench(s, "tool", (0.3, "push", -0.2, 0.0, 0.1, "twist"))
    M.HybridWrench(s, "elbow", ("lift", 0.4, 0.0, 0.2, 0.0, -0.1))
    M.SpatialWrench(s, "slider", (0.1, -0.3, "shove", 0.05, "spin", 0.0))
    return s


def fourbar() -> M.System:
    """Planar loop closed by a PointToPoint2D = two PointToPoint1D (trep/constraints/point.py,
    trep/_trep/constraints/point.c:16-55), one torque input.  nd 3, nu 1, nc 2."""
    s = M.System(name="fourbar")
    s.import_frames([
        M.rx("a1"), [M.tz(-1.0, mass=1.0, name="A1"), [M.rx("a2"), [M.tz(-1.2, name="tipA", mass=0.7)]]],
        M.ty(1.5), [M.rx("b1"), [M.tz(-1.4, name="tipB", mass=1.3)]]])
    M.Gravity(s, (0, 0, -9.8))
    M.Damping(s, 0.05)
    M.ConfigForce(s, "a1", "torque")
    M.PointToPoint2D(s, "yz", "tipA", "tipB")
    return s


def loop3d() -> M.System:
    """Spatial loop (joints about all three axes) closed by a PointToPoint3D.  nd 6, nc 3."""
    s = M.System(name="loop3d")
    s.import_frames([
        M.rz("a1"), [M.ry("a2"), [M.tx(1.0, mass=1.0), [M.rx("a3"), [M.tz(-0.8, name="tipA", mass=0.6)]]]],
        M.tx(1.2), [M.ry("b1"), [M.rz("b2"), [M.ty(0.9, mass=0.8), [M.rx("b3"), [M.tz(-0.5, name="tipB", mass=0.4)]]]]]])
    M.Gravity(s, (0, 0, -9.8))
    M.Damping(s, 0.1)
    M.PointToPoint3D(s, "tipA", "tipB")
    return s


def rod() -> M.System:
    """Two pendulums joined by a rigid rod: Distance with a fixed length (trep/constraints/distance.py,
    trep/_trep/constraints/distance.c:16-136 with config == NULL); the second pivot rides on a kinematic
    slide.  nd 3, nk 1, nc 1."""
    s = M.System(name="rod")
    s.import_frames([
        M.rx("th1"), [M.tz(-1.0, name="m1", mass=1.0)],
        M.tx("slide", kinematic=True), [M.ty(1.0), [M.ry("th2"), [M.rz("th3"), [M.tz(-1.5, name="m2", mass=2.0)]]]]])
    M.Gravity(s, (0, 0, -9.8))
    M.Damping(s, 0.02)
    M.Distance(s, "m1", "m2", 1.25)
    return s


def spring_arms() -> M.System:
    """Two spatial arms joined by LinearSprings (tip to tip, and elbow to a point fixed in the world), the second
    arm's base on a kinematic slide, one torque input, and a distance constraint between the mid links: a
    mid-size system with springs for the cooperative kernels (potentials/linearspring.c:30-74).  nd 6, nk 1,
    nu 1, nc 1."""
    s = M.System(name="spring_arms")
    s.import_frames([
        M.tx(0.3), [M.tz(1.0, name="anchor")],
        M.rz("a1"), [M.ry("a2"), [M.tx(1.0, mass=1.0, name="elbowA"), [M.rx("a3"), [M.tz(-0.8, name="tipA", mass=0.6)]]]],
        M.tx("slide", kinematic=True), [M.ty(1.1), [M.ry("b1"), [M.rz("b2"), [M.ty(0.9, mass=0.8, name="elbowB"),
                                                   [M.rx("b3"), [M.tz(-0.5, name="tipB", mass=0.4)]]]]]]])
    M.Gravity(s, (0, 0, -9.8))
    M.Damping(s, 0.05)
    M.ConfigForce(s, "a1", "torque")
    M.LinearSpring(s, "tipA", "tipB", k=15.0, x0=0.7)
    M.LinearSpring(s, "elbowA", "anchor", k=4.0, x0=0.5)
    M.Distance(s, "elbowA", "elbowB", 1.6)
    return s


def damper_only() -> M.System:
    """examples/dual_pendulums.py without the LinearSpring: the one system on which the reference's
    _calc_deriv2 reaches the LinearDamper second derivatives (forces/lineardamper.c:60-107)."""
    s = M.System(name="damper_only")
    s.import_frames([
        M.rx("theta1"), [M.tz(2, mass=1, name="pend1")],
        M.ty(1), [M.rx("theta2"), [M.tz(2, mass=1, name="pend2")]]])
    M.LinearDamper(s, "pend1", "pend2", c=1)
    M.Gravity(s, name="Gravity")
    return s


SPLINE_DATA = [(-2.0, -3.0), (-0.5, -0.4, 1.0), (0.0, 0.0), (0.7, 0.9), (2.0, 1.5, 0.2)]


def spline_pendulum() -> M.System:
    """Two-link pendulum with a spline-defined joint spring on the first joint and a linear one on
    the second (trep/potentials/nonlinear_config_spring.py; spline data in the style of
    examples/spline.py).  The quintic coefficients are the ones the reference's trep.Spline fits to
    SPLINE_DATA on the host; they are stored in data/spline_pendulum.json by oracle/gen_golden_f4.py."""
    import json
    with open(os.path.join(_DATA, "spline_pendulum.json")) as fh:
        tab = json.load(fh)
    s = M.System(name="spline_pendulum")
    s.import_frames([
        M.rx("theta1"), [M.tz(-1.0, mass=1.0), [
            M.ry("theta2"), [M.tz(-0.6, mass=0.5)]]]])
    M.Gravity(s, (0, 0, -9.8))
    M.NonlinearConfigSpring(s, "theta1", M.Spline(tab["x_points"], tab["coefficients"]), m=1.5, b=0.1)
    M.ConfigSpring(s, "theta2", k=2.0, q0=0.3)
    M.Damping(s, 0.02)
    return s


def puppet_desc() -> SystemDesc:
    """Marionette with string constraints (trep/puppets/puppets.py:220-310,
    examples/puppet-optimization.py:217-218): nd=22, nk=18, nc=6, 86 frames."""
    return SystemDesc.load(os.path.join(_DATA, "puppet.json"))


def named_desc(name) -> SystemDesc:
    if name == "puppet":
        return puppet_desc()
    table = {
        "pendulum1": lambda: pendulum(1), "pendulum5": lambda: pendulum(5),
        "damped_pendulum": damped_pendulum, "pend_on_cart1": lambda: pend_on_cart(False),
        "pend_on_cart2": lambda: pend_on_cart(True), "dual_pendulums": dual_pendulums,
        "tase_pendulum": tase_pendulum, "pccd": pccd, "wrench_arm": wrench_arm,
        "spline_pendulum": spline_pendulum, "fourbar": fourbar, "loop3d": loop3d, "rod": rod,
        "damper_only": damper_only, "spring_arms": spring_arms,
    }
    return table[name]().describe()


NAMED = ["pendulum1", "pendulum5", "damped_pendulum", "pend_on_cart1", "pend_on_cart2",
         "dual_pendulums", "tase_pendulum", "puppet"]
# systems exercising the plugin kinds beyond BASELINE.json's configs
EXTRA = ["pccd", "wrench_arm", "spline_pendulum"]
# parity systems for the constraint / force kinds of SURVEY 8a rows a7, a9 that BASELINE.json's configs do not
# exercise: PointToPoint1D/2D/3D, fixed-length Distance, LinearDamper second derivatives
PARITY = ["fourbar", "loop3d", "rod", "damper_only"]
# a mid-size system with LinearSprings that the cooperative kernels run (round 2)
PARITY_SPRING = ["spring_arms"]
