"""Host-side model-building mirror of the slice of trep's Python API that the hot path needs.

Only what is required to *describe* a system for the batched MidpointVI path lives here:
frame trees (``tx/ty/tz/rx/ry/rz/const_se3/const_txyz`` + ``System.import_frames``,
reference ``trep/frame.py:21-37,201-220``), configs/inputs, and the built-in plugin kinds of
SURVEY.md §8 (Gravity, LinearSpring, ConfigSpring, Damping, ConfigForce, LinearDamper,
Distance, PointToPoint1D/2D/3D).  Names and argument meaning follow the reference so that
model scripts read the same; evaluation does NOT happen here — ``System.describe()`` flattens
the model to a :class:`trep_b200.desc.SystemDesc` for the CUDA library.

``flatten_trep_system`` does the same for a live reference ``trep.System`` (duck-typed on the
attribute names of SURVEY.md Appendix B) — that is the drop-in route.
"""
from __future__ import annotations

import numpy as np

from . import desc as D


class FrameDef:
    def __init__(self, kind, param, name, kinematic, mass):
        self.kind, self.param, self.name, self.kinematic, self.mass = kind, param, name, kinematic, mass


def tx(param, name=None, kinematic=False, mass=0.0):
    return FrameDef(D.TX, param, name, kinematic, mass)


def ty(param, name=None, kinematic=False, mass=0.0):
    return FrameDef(D.TY, param, name, kinematic, mass)


def tz(param, name=None, kinematic=False, mass=0.0):
    return FrameDef(D.TZ, param, name, kinematic, mass)


def rx(param, name=None, kinematic=False, mass=0.0):
    return FrameDef(D.RX, param, name, kinematic, mass)


def ry(param, name=None, kinematic=False, mass=0.0):
    return FrameDef(D.RY, param, name, kinematic, mass)


def rz(param, name=None, kinematic=False, mass=0.0):
    return FrameDef(D.RZ, param, name, kinematic, mass)


def const_se3(se3, name=None, kinematic=False, mass=0.0):
    return FrameDef(D.CONST_SE3, se3, name, kinematic, mass)


def const_txyz(xyz, name=None, kinematic=False, mass=0.0):
    return FrameDef(D.CONST_SE3, [[1, 0, 0], [0, 1, 0], [0, 0, 1], xyz], name, kinematic, mass)


class Config:
    def __init__(self, system, name, kinematic=False):
        self.system, self.name, self.kinematic = system, name, bool(kinematic)
        self.q = 0.0
        self.frame = None
        (system.kin_configs if kinematic else system.dyn_configs).append(self)

    @property
    def index(self):
        return self.system.configs.index(self)


class Input:
    def __init__(self, system, name):
        self.system, self.name = system, name
        system.inputs.append(self)

    @property
    def index(self):
        return self.system.inputs.index(self)


class Frame:
    def __init__(self, parent, transform, param, name=None, kinematic=False, mass=0.0):
        self.name = name
        self.children = []
        self.config = None
        self.value = 0.0
        self.se3 = np.hstack([np.eye(3), np.zeros((3, 1))])
        self.transform = transform
        if transform == D.WORLD:
            self.system, self.parent = parent, None
        else:
            self.system, self.parent = parent.system, parent
            parent.children.append(self)
            if transform == D.CONST_SE3:
                # param = (x-axis, y-axis, z-axis, position)  (trep/frame.py set_SE3)
                m = np.array(param, dtype=float)
                self.se3 = np.hstack([m[0:3].T, m[3].reshape(3, 1)])
            elif isinstance(param, str):
                self.config = Config(self.system, param, kinematic=kinematic)
                self.config.frame = self
            else:
                self.value = float(param)
        self.set_mass(mass)

    def set_mass(self, mass, Ixx=0.0, Iyy=0.0, Izz=0.0):
        if isinstance(mass, (list, tuple, np.ndarray)):
            mass, Ixx, Iyy, Izz = [float(v) for v in mass]
        self.mass = (float(mass), float(Ixx), float(Iyy), float(Izz))

    def import_frames(self, children):
        children = list(children)
        while children:
            info = children.pop(0)
            if not isinstance(info, FrameDef):
                raise TypeError("Frame definition expected instead of: %r" % (info,))
            frame = Frame(self, info.kind, info.param, name=info.name,
                          kinematic=info.kinematic, mass=info.mass)
            if children and isinstance(children[0], list):
                frame.import_frames(children.pop(0))

    def flatten_tree(self):
        out = [self]
        for c in self.children:
            out += c.flatten_tree()
        return out


class System:
    """Mirror of ``trep.System`` restricted to structure (no evaluation)."""

    def __init__(self, name=""):
        self.name = name
        self.dyn_configs, self.kin_configs, self.inputs = [], [], []
        self.potentials, self.forces, self.constraints = [], [], []
        self.world_frame = Frame(self, D.WORLD, None, name="World")

    # -- structure ------------------------------------------------------------------------
    @property
    def configs(self):
        return self.dyn_configs + self.kin_configs

    @property
    def frames(self):
        return self.world_frame.flatten_tree()

    nQ = property(lambda s: len(s.configs))
    nQd = property(lambda s: len(s.dyn_configs))
    nQk = property(lambda s: len(s.kin_configs))
    nu = property(lambda s: len(s.inputs))
    nc = property(lambda s: len(s.constraints))

    def import_frames(self, children):
        self.world_frame.import_frames(children)

    def get_frame(self, ident):
        if isinstance(ident, Frame):
            return ident
        for f in self.frames:
            if f.name == ident:
                return f
        return None

    def get_config(self, ident):
        if isinstance(ident, Config):
            return ident
        for c in self.configs:
            if c.name == ident:
                return c
        return None

    def get_input(self, ident):
        if isinstance(ident, Input):
            return ident
        for u in self.inputs:
            if u.name == ident:
                return u
        return None

    @property
    def q(self):
        return np.array([c.q for c in self.configs])

    @q.setter
    def q(self, value):
        if isinstance(value, dict):
            for k, v in value.items():
                self.get_config(k).q = float(v)
        else:
            for c, v in zip(self.configs, np.atleast_1d(value)):
                c.q = float(v)

    # -- flattening -----------------------------------------------------------------------
    def describe(self) -> D.SystemDesc:
        frames = self.frames
        fidx = {id(f): i for i, f in enumerate(frames)}
        configs = self.configs
        cidx = {id(c): i for i, c in enumerate(configs)}
        ipool, dpool = [], []
        pk, pi, pd = [], [], []
        for p in self.potentials:
            k, ii, dd = p._record(fidx, cidx, ipool, dpool)
            pk.append(k); pi.append(ii); pd.append(dd)
        fk, fi, fd = [], [], []
        for f in self.forces:
            k, ii, dd = f._record(fidx, cidx, ipool, dpool)
            fk.append(k); fi.append(ii); fd.append(dd)
        ck, ci, cd = [], [], []
        for c in self.constraints:
            k, ii, dd = c._record(fidx, cidx, ipool, dpool)
            ck.append(k); ci.append(ii); cd.append(dd)
        return D.make_desc(
            frame_parent=[-1 if f.parent is None else fidx[id(f.parent)] for f in frames],
            frame_kind=[f.transform for f in frames],
            frame_config=[-1 if f.config is None else cidx[id(f.config)] for f in frames],
            frame_value=[f.value for f in frames],
            frame_se3=[f.se3.reshape(-1) for f in frames],
            frame_mass=[f.mass for f in frames],
            nd=self.nQd, nk=self.nQk, nu=self.nu,
            pot_kind=pk, pot_i=pi, pot_d=pd, force_kind=fk, force_i=fi, force_d=fd,
            con_kind=ck, con_i=ci, con_d=cd, ipool=ipool, dpool=dpool,
            frame_names=[f.name or "" for f in frames],
            config_names=[c.name for c in configs],
            input_names=[u.name for u in self.inputs], name=self.name)


def _pad4(x, fill):
    x = list(x)
    return x + [fill] * (4 - len(x))


# ---- potentials (trep/potentials/*.py) -----------------------------------------------------
class Gravity:
    def __init__(self, system, gravity=(0.0, 0.0, -9.8), name=None):
        self.system, self.gravity, self.name = system, tuple(float(g) for g in gravity), name
        system.potentials.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        return D.POT_GRAVITY, _pad4([], -1), _pad4(self.gravity, 0.0)


class LinearSpring:
    def __init__(self, system, frame1, frame2, k, x0=0, name=None):
        self.system, self.k, self.x0, self.name = system, float(k), float(x0), name
        self.frame1, self.frame2 = system.get_frame(frame1), system.get_frame(frame2)
        if self.frame1 is None:
            raise ValueError("Could not find frame %r" % (frame1,))
        if self.frame2 is None:
            raise ValueError("Could not find frame %r" % (frame2,))
        system.potentials.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        return (D.POT_LINEAR_SPRING, _pad4([fidx[id(self.frame1)], fidx[id(self.frame2)]], -1),
                _pad4([self.k, self.x0], 0.0))


class ConfigSpring:
    def __init__(self, system, config, k, q0=0.0, name=None):
        self.system, self.k, self.q0, self.name = system, float(k), float(q0), name
        self.config = system.get_config(config)
        if self.config is None:
            raise ValueError("Could not find config %r" % (config,))
        system.potentials.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        return D.POT_CONFIG_SPRING, _pad4([cidx[id(self.config)]], -1), _pad4([self.k, self.q0], 0.0)


class Spline:
    """Piecewise quintic y(x) as the reference evaluates it (trep/_trep/spline.c:7-62): x points and,
    per segment, six coefficients (highest power first) of dx = x - x_i.  The reference fits the
    coefficients on the host (trep/spline.py, numpy.linalg.solve); the batched path only evaluates,
    so this mirror takes the fitted tables - from a reference ``trep.Spline`` via
    ``Spline.from_reference`` or from stored arrays."""
    def __init__(self, x_points, coefficients):
        self.x_points = np.ascontiguousarray(x_points, dtype=float)
        self.coefficients = np.ascontiguousarray(coefficients, dtype=float).reshape(-1, 6)
        if self.coefficients.shape[0] != self.x_points.shape[0] - 1:
            raise ValueError("a spline over n x points has n-1 coefficient rows")

    @staticmethod
    def from_reference(spline):
        return Spline(np.array(spline._x_points), np.array(spline._coefficients))


class NonlinearConfigSpring:
    """dV/dq = -spline(m q + b) (trep/potentials/nonlinear_config_spring.py)."""
    def __init__(self, system, config, spline, m=1.0, b=0.0, name=None):
        self.system, self.name = system, name
        self.config = system.get_config(config)
        if self.config is None:
            raise ValueError("Could not find config %r" % (config,))
        self.spline, self.m, self.b = spline, float(m), float(b)
        system.potentials.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        off = len(dpool)
        dpool.extend(self.spline.x_points.tolist())
        dpool.extend(self.spline.coefficients.reshape(-1).tolist())
        return (D.POT_NONLINEAR_CONFIG_SPRING, [cidx[id(self.config)], off, len(self.spline.x_points), -1],
                _pad4([self.m, self.b], 0.0))


# ---- forces (trep/forces/*.py) -------------------------------------------------------------
class Damping:
    def __init__(self, system, default=0.0, coefficients={}, name=None):
        self.system, self.default, self.name = system, float(default), name
        self.coefficients = {system.get_config(c): float(v) for c, v in coefficients.items()}
        system.forces.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        nd = self.system.nQd
        coeff = [self.default] * nd
        for c, v in self.coefficients.items():
            if cidx[id(c)] < nd:
                coeff[cidx[id(c)]] = v
        off = len(dpool)
        dpool.extend(coeff)
        return D.FORCE_DAMPING, _pad4([off, nd], -1), _pad4([self.default], 0.0)


class ConfigForce:
    def __init__(self, system, config, finput, name=None):
        self.system, self.name = system, name
        self.config = system.get_config(config)
        if self.config is None:
            raise ValueError("Could not find config %r" % (config,))
        self.input = finput if isinstance(finput, Input) else Input(system, finput)
        system.forces.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        return (D.FORCE_CONFIG, _pad4([cidx[id(self.config)], self.input.index], -1), _pad4([], 0.0))


class LinearDamper:
    def __init__(self, system, frame1, frame2, c, name=None):
        self.system, self.c, self.name = system, float(c), name
        self.frame1, self.frame2 = system.get_frame(frame1), system.get_frame(frame2)
        if self.frame1 is None:
            raise ValueError("Could not find frame %r" % (frame1,))
        if self.frame2 is None:
            raise ValueError("Could not find frame %r" % (frame2,))
        system.forces.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        off = len(ipool)
        ipool.extend([fidx[id(self.frame1)], fidx[id(self.frame2)]])
        return D.FORCE_LINEAR_DAMPER, _pad4([off, 2], -1), _pad4([self.c], 0.0)


class _Wrench:
    """Wrench on a frame; each of the six components is a constant or the name of an input
    (trep/forces/bodywrench.py, hybridwrench.py, spatialwrench.py)."""
    KIND = None

    def __init__(self, system, frame, wrench=tuple(), name=None):
        self.system, self.name = system, name
        self.frame = system.get_frame(frame)
        if self.frame is None:
            raise ValueError("Could not find frame %r" % (frame,))
        wrench = (list(wrench) + [0.0] * 6)[:6]
        self.inputs, self.constants = [None] * 6, [0.0] * 6
        for i in range(6):
            if isinstance(wrench[i], str):
                self.inputs[i] = Input(system, wrench[i])
            else:
                self.constants[i] = float(wrench[i])
        system.forces.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        io, do = len(ipool), len(dpool)
        ipool.extend([-1 if u is None else u.index for u in self.inputs])
        dpool.extend(self.constants)
        return self.KIND, [fidx[id(self.frame)], io, do, -1], _pad4([], 0.0)


class BodyWrench(_Wrench):
    KIND = D.FORCE_BODY_WRENCH


class HybridWrench(_Wrench):
    KIND = D.FORCE_HYBRID_WRENCH


class SpatialWrench(_Wrench):
    KIND = D.FORCE_SPATIAL_WRENCH


# ---- constraints (trep/constraints/*.py) ---------------------------------------------------
class Distance:
    def __init__(self, system, frame1, frame2, distance, name=None, tolerance=1e-10):
        self.system, self.name, self.tolerance = system, name, float(tolerance)
        self.frame1, self.frame2 = system.get_frame(frame1), system.get_frame(frame2)
        assert self.frame1 is not None and self.frame2 is not None
        if isinstance(distance, str):
            self.config, self.distance = Config(system, distance, kinematic=True), 0.0
        else:
            self.config, self.distance = None, float(distance)
        system.constraints.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        c = -1 if self.config is None else cidx[id(self.config)]
        return (D.CON_DISTANCE, [fidx[id(self.frame1)], fidx[id(self.frame2)], c, -1],
                _pad4([self.distance, self.tolerance], 0.0))


class PointToPoint1D:
    def __init__(self, system, axis, frame1, frame2, name=None, tolerance=1e-10):
        self.system, self.name, self.tolerance = system, name, float(tolerance)
        self.frame1, self.frame2 = system.get_frame(frame1), system.get_frame(frame2)
        assert self.frame1 is not None and self.frame2 is not None
        self.component = {"x": 0, "y": 1, "z": 2, "X": 0, "Y": 1, "Z": 2}[axis]
        system.constraints.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        return (D.CON_POINT1D, [fidx[id(self.frame1)], fidx[id(self.frame2)], self.component, -1],
                _pad4([0.0, self.tolerance], 0.0))


class PointOnPlane:
    """Point frame stays on the plane through plane_frame's origin with the given normal (fixed in
    plane_frame) (trep/constraints/plane.py)."""
    def __init__(self, system, plane_frame, plane_normal, point_frame, name=None, tolerance=1e-10):
        self.system, self.name, self.tolerance = system, name, float(tolerance)
        self.plane_frame, self.point_frame = system.get_frame(plane_frame), system.get_frame(point_frame)
        assert self.plane_frame is not None and self.point_frame is not None
        self.normal = tuple(float(x) for x in plane_normal)
        system.constraints.append(self)

    def _record(self, fidx, cidx, ipool, dpool):
        n = self.normal
        return (D.CON_PLANE, [fidx[id(self.plane_frame)], fidx[id(self.point_frame)], -1, -1],
                [n[0], self.tolerance, n[1], n[2]])


class PointToPoint2D:
    def __init__(self, system, plane, frame1, frame2, name=None):
        axes = {0: "yz", 1: "xz", 2: "xy"}[{"yz": 0, "zy": 0, "xz": 1, "zx": 1, "xy": 2, "yx": 2}[plane.lower()]]
        for a in axes:
            PointToPoint1D(system, a, frame1, frame2, name)


class PointToPoint3D:
    def __init__(self, system, frame1, frame2, name=None):
        for a in "xyz":
            PointToPoint1D(system, a, frame1, frame2, name)


class _NS:
    pass


potentials = _NS()
potentials.Gravity, potentials.LinearSpring, potentials.ConfigSpring = Gravity, LinearSpring, ConfigSpring
potentials.NonlinearConfigSpring = NonlinearConfigSpring
forces = _NS()
forces.Damping, forces.ConfigForce, forces.LinearDamper = Damping, ConfigForce, LinearDamper
forces.BodyWrench, forces.HybridWrench, forces.SpatialWrench = BodyWrench, HybridWrench, SpatialWrench
constraints = _NS()
constraints.Distance, constraints.PointToPoint1D = Distance, PointToPoint1D
constraints.PointToPoint2D, constraints.PointToPoint3D = PointToPoint2D, PointToPoint3D
constraints.PointOnPlane = PointOnPlane


# ------------------------------------------------------------------------------------------
# Flatten a live reference `trep.System` (attribute names: SURVEY.md Appendix B)
# ------------------------------------------------------------------------------------------
def flatten_trep_system(system, name="") -> D.SystemDesc:
    """Build a SystemDesc from a live reference ``trep.System`` without importing trep.

    Raises ``TypeError`` for any potential/force/constraint kind that has no device
    implementation (Python-defined plugins): the batched
    path refuses rather than falls back.
    """
    frames = list(system.frames)
    fidx = {id(f): i for i, f in enumerate(frames)}
    configs = list(system.configs)
    cidx = {id(c): i for i, c in enumerate(configs)}
    inputs = list(system.inputs)
    uidx = {id(u): i for i, u in enumerate(inputs)}
    kind_of = {"WORLD": D.WORLD, "TX": D.TX, "TY": D.TY, "TZ": D.TZ, "RX": D.RX, "RY": D.RY,
               "RZ": D.RZ, "CONST_SE3": D.CONST_SE3}

    def tname(f):
        t = f.transform_type
        s = getattr(t, "name", None) or str(t)
        for k in kind_of:
            if s.upper().endswith(k) or s.upper() == k:
                return k
        raise TypeError("unknown frame transform %r" % (t,))

    parent, kind, config, value, se3, mass, names = [], [], [], [], [], [], []
    for f in frames:
        k = kind_of[tname(f)]
        parent.append(-1 if f.parent is None else fidx[id(f.parent)])
        kind.append(k)
        config.append(-1 if f.config is None else cidx[id(f.config)])
        value.append(0.0 if (f.config is not None or k in (D.WORLD, D.CONST_SE3)) else float(f.transform_value))
        if k == D.CONST_SE3:
            se3.append(np.array(f.lg())[:3, :].reshape(-1))
        else:
            se3.append(np.hstack([np.eye(3), np.zeros((3, 1))]).reshape(-1))
        mass.append([float(f.mass), float(f.Ixx), float(f.Iyy), float(f.Izz)])
        names.append(f.name or "")

    ipool, dpool = [], []
    pk, pi, pd = [], [], []
    for p in system.potentials:
        cls = type(p).__name__
        if cls == "Gravity":
            pk.append(D.POT_GRAVITY); pi.append([-1] * 4); pd.append(list(map(float, p.gravity)) + [0.0])
        elif cls == "LinearSpring":
            pk.append(D.POT_LINEAR_SPRING)
            pi.append([fidx[id(p.frame1)], fidx[id(p.frame2)], -1, -1]); pd.append([float(p.k), float(p.x0), 0, 0])
        elif cls == "ConfigSpring":
            pk.append(D.POT_CONFIG_SPRING)
            pi.append([cidx[id(p.config)], -1, -1, -1]); pd.append([float(p.k), float(p.q0), 0, 0])
        elif cls == "NonlinearConfigSpring":
            off = len(dpool)
            xs = np.array(p.spline._x_points, dtype=float)
            dpool.extend(xs.tolist())
            dpool.extend(np.array(p.spline._coefficients, dtype=float).reshape(-1).tolist())
            pk.append(D.POT_NONLINEAR_CONFIG_SPRING)
            pi.append([cidx[id(p.config)], off, len(xs), -1]); pd.append([float(p._m), float(p._b), 0, 0])
        else:
            raise TypeError("potential %s has no batched device implementation" % cls)
    fk, fi, fd = [], [], []
    for f in system.forces:
        cls = type(f).__name__
        if cls == "Damping":
            off = len(dpool)
            dpool.extend(np.array(f._coefficients, dtype=float).tolist())
            fk.append(D.FORCE_DAMPING); fi.append([off, system.nQd, -1, -1]); fd.append([float(f._default), 0, 0, 0])
        elif cls == "ConfigForce":
            fk.append(D.FORCE_CONFIG); fi.append([cidx[id(f.config)], uidx[id(f.finput)], -1, -1]); fd.append([0.0] * 4)
        elif cls == "LinearDamper":
            off = len(ipool)
            path = list(f._path._frames)
            ipool.extend([fidx[id(x)] for x in path])
            fk.append(D.FORCE_LINEAR_DAMPER); fi.append([off, len(path), -1, -1]); fd.append([float(f.c), 0, 0, 0])
        elif cls in ("BodyWrench", "HybridWrench", "SpatialWrench"):
            io, do = len(ipool), len(dpool)
            ipool.extend([-1 if u is None else uidx[id(u)] for u in f._wrench_vars])
            dpool.extend([float(x) for x in f._wrench_cons])
            wkind = {"BodyWrench": D.FORCE_BODY_WRENCH, "HybridWrench": D.FORCE_HYBRID_WRENCH,
                     "SpatialWrench": D.FORCE_SPATIAL_WRENCH}[cls]
            fk.append(wkind); fi.append([fidx[id(f._frame)], io, do, -1]); fd.append([0.0] * 4)
        else:
            raise TypeError("force %s has no batched device implementation" % cls)
    ck, ci, cd = [], [], []
    for c in system.constraints:
        cls = type(c).__name__
        if cls == "Distance":
            cfg = -1 if c.config is None else cidx[id(c.config)]
            dist = 0.0 if c.config is not None else float(c._distance)
            ck.append(D.CON_DISTANCE); ci.append([fidx[id(c.frame1)], fidx[id(c.frame2)], cfg, -1])
            cd.append([dist, float(c.tolerance), 0, 0])
        elif cls == "PointToPoint1D":
            ck.append(D.CON_POINT1D); ci.append([fidx[id(c.frame1)], fidx[id(c.frame2)], int(c._component), -1])
            cd.append([0.0, float(c.tolerance), 0, 0])
        elif cls == "PointOnPlane":
            n = [float(x) for x in c.normal]
            ck.append(D.CON_PLANE); ci.append([fidx[id(c.plane_frame)], fidx[id(c.point_frame)], -1, -1])
            cd.append([n[0], float(c.tolerance), n[1], n[2]])
        else:
            raise TypeError("constraint %s has no batched device implementation" % cls)
    return D.make_desc(parent, kind, config, value, se3, mass, system.nQd, system.nQk, system.nu,
                       pk, pi, pd, fk, fi, fd, ck, ci, cd, ipool, dpool, names,
                       [c.name for c in configs], [u.name for u in inputs], name)
