"""Reference-side system builders + a batch driver over the reference's own MidpointVI.

TEST INFRASTRUCTURE ONLY (the oracle): imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by trep_b200/.

Everything here runs the UNMODIFIED-NUMERICS reference built into oracle/_ref by
oracle/build_ref.py (on the GPU box the prebuilt oracle/_ref travels with the snapshot;
/root/reference is not needed at run time).
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import build_ref  # noqa: E402

if os.path.isdir("/root/reference/trep"):
    build_ref.build("/root/reference")
trep = build_ref.import_ref()
from trep import tx, ty, tz, rx, ry, rz  # noqa: E402,F401
from trep import discopt  # noqa: E402,F401
import trep.puppets  # noqa: E402


# ---- reference-side builders (same scripts as the examples cited in trep_b200/systems.py) ----
def ref_pendulum(links):
    system = trep.System()
    trep.potentials.Gravity(system, name="Gravity")
    frame = system.world_frame
    for link in range(links):
        frame = trep.Frame(frame, trep.RX, "link-%d" % link, "link-%d" % link)
        frame = trep.Frame(frame, trep.TZ, -1)
        frame.set_mass(1.0)
    system.get_config("link-0").q = math.pi / 4.0
    return system


def ref_damped_pendulum():
    system = trep.System()
    system.import_frames([ty(3), rx("theta"), [tz(-3, mass=1)]])
    trep.potentials.Gravity(system, (0, 0, -9.8))
    trep.forces.Damping(system, 1.2)
    return system


def ref_pend_on_cart(torque):
    system = trep.System()
    system.import_frames([
        tx('x', name='Cart', mass=10.0), [
            rz('theta', name="PendulumBase"), [
                ty(-1.0, name="Pendulum", mass=1.0)]]])
    trep.potentials.Gravity(system, (0, -9.8, 0))
    trep.forces.Damping(system, 0.01)
    trep.forces.ConfigForce(system, 'x', 'x-force')
    if torque:
        trep.forces.ConfigForce(system, 'theta', 'theta-force')
    return system


def ref_dual_pendulums():
    system = trep.System()
    system.import_frames([
        rx('theta1'), [tz(2, mass=1, name='pend1')],
        ty(1), [rx('theta2'), [tz(2, mass=1, name='pend2')]]])
    trep.potentials.LinearSpring(system, 'pend1', 'pend2', k=20, x0=1)
    trep.forces.LinearDamper(system, 'pend1', 'pend2', c=1)
    trep.potentials.Gravity(system, name="Gravity")
    system.q = [3, -3]
    return system


def ref_tase_pendulum():
    system = trep.System()
    system.import_frames([trep.rz("theta_1", name="PendAngle"), [trep.ty(-1.0, name="PendMass", mass=1.0)]])
    trep.potentials.Gravity(system, (0, -9.8, 0))
    trep.forces.ConfigForce(system, "theta_1", "tau")
    return system


def ref_pccd():
    """examples/pccd.py:20-66 (PointOnPlane constraints)."""
    system = trep.System()
    system.import_frames([
        rx('J', name='J'), [
            tz(-0.5, name='I', mass=1),
            tz(-1), [
                rx('H', name='H'), [
                    tz(-1, name='G', mass=1),
                    tz(-2, name='O2')]]],
        ty(1.5), [
            rx('K', name='K'), [
                tz(-1, name='L', mass=1),
                tz(-2), [
                    rx('M', name='M'), [
                        tz(-0.5, name='N', mass=1),
                        tz(-1.0, name='O')]]]],
        ty(-1.5), [
            rx('A', name='A'), [
                tz(-1, name='B', mass=1),
                tz(-2), [
                    rx('C', name='C'), [
                        tz(-0.375, name='D', mass=1),
                        tz(-0.75), [
                            rx('E', name='E'), [
                                tz(-0.5, name='F', mass=1),
                                tz(-1.0, name='G2')]]]]]]])
    trep.potentials.Gravity(system, (0, 0, -9.8))
    trep.forces.Damping(system, 0.1)
    trep.constraints.PointOnPlane(system, 'O', (0, 1, 0), 'O2')
    trep.constraints.PointOnPlane(system, 'O', (0, 0, 1), 'O2')
    trep.constraints.PointOnPlane(system, 'G', (0, 1, 0), 'G2')
    trep.constraints.PointOnPlane(system, 'G', (0, 0, 1), 'G2')
    system.q = {'K': 0.523599, 'M': -1.34537, 'J': -0.523599, 'H': 1.21009, 'A': -0.523599,
                'C': 1.5385, 'E': 1.22497}
    system.satisfy_constraints()
    return system


def ref_wrench_arm():
    """Same script as trep_b200/systems.py:wrench_arm with the reference's own classes."""
    system = trep.System()
    system.import_frames([
        rz('yaw', name='base'), [
            tz(0.4, name='shoulder', mass=2.0), [
                ry('pitch', name='upper'), [
                    tx(0.7, name='elbow', mass=1.5), [
                        tx('reach', name='slider'), [
                            trep.const_txyz((0.1, -0.2, 0.3), name='tool', mass=0.5)]]]]]])
    system.get_frame('shoulder').set_mass(2.0, 0.1, 0.2, 0.3)
    system.get_frame('tool').set_mass(0.5, 0.02, 0.03, 0.01)
    trep.potentials.Gravity(system, (0, 0, -9.8))
    trep.forces.Damping(system, 0.05)
    trep.forces.BodyWrench(system, 'tool', (0.3, 'push', -0.2, 0.0, 0.1, 'twist'))
    trep.forces.HybridWrench(system, 'elbow', ('lift', 0.4, 0.0, 0.2, 0.0, -0.1))
    trep.forces.SpatialWrench(system, 'slider', (0.1, -0.3, 'shove', 0.05, 'spin', 0.0))
    return system


def ref_spline_pendulum():
    """Same script as trep_b200/systems.py:spline_pendulum with the reference's own classes."""
    from trep_b200.systems import SPLINE_DATA
    system = trep.System()
    system.import_frames([
        rx('theta1'), [tz(-1.0, mass=1.0), [
            ry('theta2'), [tz(-0.6, mass=0.5)]]]])
    trep.potentials.Gravity(system, (0, 0, -9.8))
    trep.potentials.NonlinearConfigSpring(system, 'theta1', trep.Spline(SPLINE_DATA), m=1.5, b=0.1)
    trep.potentials.ConfigSpring(system, 'theta2', k=2.0, q0=0.3)
    trep.forces.Damping(system, 0.02)
    return system


def ref_fourbar():
    """Planar loop closed by a PointToPoint2D (two PointToPoint1D, constraints/point.c:16-55); one input."""
    system = trep.System()
    system.import_frames([
        rx('a1'), [tz(-1.0, mass=1.0, name='A1'), [rx('a2'), [tz(-1.2, name='tipA', mass=0.7)]]],
        ty(1.5), [rx('b1'), [tz(-1.4, name='tipB', mass=1.3)]]])
    trep.potentials.Gravity(system, (0, 0, -9.8))
    trep.forces.Damping(system, 0.05)
    trep.forces.ConfigForce(system, 'a1', 'torque')
    trep.constraints.PointToPoint2D(system, 'yz', 'tipA', 'tipB')
    system.q = {'a1': 0.4, 'a2': 0.9, 'b1': -0.3}
    system.satisfy_constraints()
    return system


def ref_loop3d():
    """Spatial loop (joints about all three axes) closed by a PointToPoint3D."""
    system = trep.System()
    system.import_frames([
        rz('a1'), [ry('a2'), [tx(1.0, mass=1.0), [rx('a3'), [tz(-0.8, name='tipA', mass=0.6)]]]],
        tx(1.2), [ry('b1'), [rz('b2'), [ty(0.9, mass=0.8), [rx('b3'), [tz(-0.5, name='tipB', mass=0.4)]]]]]])
    trep.potentials.Gravity(system, (0, 0, -9.8))
    trep.forces.Damping(system, 0.1)
    trep.constraints.PointToPoint3D(system, 'tipA', 'tipB')
    system.q = {'a1': 0.3, 'a2': -0.2, 'a3': 0.5, 'b1': 0.4, 'b2': 0.8, 'b3': -0.6}
    system.satisfy_constraints()
    return system


def ref_rod():
    """Two pendulums joined by a rigid rod: Distance with a FIXED length (constraints/distance.c:16-136,
    the config == NULL branch); the second pivot rides on a kinematic slide."""
    system = trep.System()
    system.import_frames([
        rx('th1'), [tz(-1.0, name='m1', mass=1.0)],
        tx('slide', kinematic=True), [ty(1.0), [ry('th2'), [rz('th3'), [tz(-1.5, name='m2', mass=2.0)]]]]])
    trep.potentials.Gravity(system, (0, 0, -9.8))
    trep.forces.Damping(system, 0.02)
    trep.constraints.Distance(system, 'm1', 'm2', 1.25)
    system.q = {'th1': 0.3, 'th2': 0.2, 'th3': -0.1}
    system.satisfy_constraints()
    return system


def ref_spring_arms():
    """Same script as trep_b200/systems.py:spring_arms with the reference's own classes."""
    system = trep.System()
    system.import_frames([
        tx(0.3), [tz(1.0, name='anchor')],
        rz('a1'), [ry('a2'), [tx(1.0, mass=1.0, name='elbowA'), [rx('a3'), [tz(-0.8, name='tipA', mass=0.6)]]]],
        tx('slide', kinematic=True), [ty(1.1), [ry('b1'), [rz('b2'), [ty(0.9, mass=0.8, name='elbowB'),
                                                 [rx('b3'), [tz(-0.5, name='tipB', mass=0.4)]]]]]]])
    trep.potentials.Gravity(system, (0, 0, -9.8))
    trep.forces.Damping(system, 0.05)
    trep.forces.ConfigForce(system, 'a1', 'torque')
    trep.potentials.LinearSpring(system, 'tipA', 'tipB', k=15.0, x0=0.7)
    trep.potentials.LinearSpring(system, 'elbowA', 'anchor', k=4.0, x0=0.5)
    trep.constraints.Distance(system, 'elbowA', 'elbowB', 1.6)
    system.q = {'a1': 0.3, 'a2': -0.2, 'a3': 0.5, 'b1': 0.4, 'b2': 0.6, 'b3': -0.6}
    system.satisfy_constraints()
    return system


def ref_damper_only():
    """examples/dual_pendulums.py without the LinearSpring (whose missing C V_dqdqdq stops the reference's
    _calc_deriv2): the LinearDamper's second derivatives (forces/lineardamper.c:60-107) are reachable."""
    system = trep.System()
    system.import_frames([
        rx('theta1'), [tz(2, mass=1, name='pend1')],
        ty(1), [rx('theta2'), [tz(2, mass=1, name='pend2')]]])
    trep.forces.LinearDamper(system, 'pend1', 'pend2', c=1)
    trep.potentials.Gravity(system, name="Gravity")
    system.q = [3, -3]
    return system


def ref_puppet():
    puppet = trep.puppets.Puppet(joint_forces=False, string_forces=False, string_constraints=True)
    puppet.q = {
        'torso_rx': -0.05, 'torso_tz': 0.0, 'lelbow_rx': 1.57, 'relbow_rx': 1.57,
        'lhip_rx': math.pi / 2 - 0.6, 'rhip_rx': math.pi / 2 - 0.6,
        'lknee_rx': -math.pi / 2 + 0.6, 'rknee_rx': -math.pi / 2 + 0.6}
    puppet.project_string_controls()
    return puppet


REF_BUILDERS = {
    "pendulum1": lambda: ref_pendulum(1), "pendulum5": lambda: ref_pendulum(5),
    "damped_pendulum": ref_damped_pendulum, "pend_on_cart1": lambda: ref_pend_on_cart(False),
    "pend_on_cart2": lambda: ref_pend_on_cart(True), "dual_pendulums": ref_dual_pendulums,
    "tase_pendulum": ref_tase_pendulum, "puppet": ref_puppet,
    "pccd": ref_pccd, "wrench_arm": ref_wrench_arm, "spline_pendulum": ref_spline_pendulum,
    "fourbar": ref_fourbar, "loop3d": ref_loop3d, "rod": ref_rod, "damper_only": ref_damper_only,
    "spring_arms": ref_spring_arms,
}


def make_mvi(name, tolerance=1e-10):
    """(system, MidpointVI) of the reference for a named system."""
    system = REF_BUILDERS[name]()
    return system, trep.MidpointVI(system, tolerance=tolerance, num_threads=1)


def run_cases(mvi, t1, t2, q1, p1, u1, k2, q2_guess=None, lambda_guess=None, deriv1=True):
    """Reference results for a batch of independent (q1,p1,u1,k2[,guess]) instances:
    initialize_from_state + step (+ _calc_deriv1 and the DSystem A/B blocks), one Python call
    per instance.  Arrays are [B][...]; t1/t2 scalars or [B]."""
    q1 = np.atleast_2d(q1)
    B = q1.shape[0]
    nq, nd, nu, nk, nc = mvi.nq, mvi.nd, mvi.nu, mvi.nk, mvi.nc
    t1 = np.broadcast_to(np.asarray(t1, float), (B,))
    t2 = np.broadcast_to(np.asarray(t2, float), (B,))
    out = dict(q2=np.zeros((B, nq)), p2=np.zeros((B, nd)), lambda1=np.zeros((B, nc)),
               iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32))
    if deriv1:
        out["A"] = np.zeros((B, 2 * nq, 2 * nq))
        out["B"] = np.zeros((B, 2 * nq, nu + nk))
    dsys = None
    for b in range(B):
        mvi.initialize_from_state(float(t1[b]), q1[b], p1[b])
        hint = None if q2_guess is None else np.array(q2_guess[b])
        lh = None if lambda_guess is None else np.array(lambda_guess[b])
        try:
            out["iters"][b] = mvi.step(float(t2[b]), tuple(u1[b]) if nu else tuple(),
                                       tuple(k2[b]) if nk else tuple(), q2_hint=hint, lambda1_hint=lh)
        except trep.ConvergenceError:
            out["status"][b] = -1
            continue
        out["q2"][b], out["p2"][b], out["lambda1"][b] = mvi.q2, mvi.p2, mvi.lambda1
        if deriv1:
            if dsys is None:
                dsys = discopt.DSystem(mvi, None)
            dsys._k = 0
            dsys._time = np.array([t1[b], t2[b]])
            out["A"][b] = dsys.fdx()
            out["B"][b] = dsys.fdu()
    return out


def run_rollout(mvi, q0, q1, dt, nsteps, u=None, k=None):
    """initialize_from_configs(0,q0,dt,q1) + nsteps of step(); u [nsteps][nu], k [nsteps][nk]."""
    mvi.initialize_from_configs(0.0, q0, dt, q1)
    p_init = mvi.p2
    iters = np.zeros(nsteps, np.int32)
    for s in range(nsteps):
        u1 = tuple() if u is None else tuple(u[s])
        k2 = tuple() if k is None else tuple(k[s])
        iters[s] = mvi.step(mvi.t2 + dt, u1, k2)
    return dict(p_init=p_init, q2=mvi.q2, p2=mvi.p2, lambda1=mvi.lambda1, iters=iters)
