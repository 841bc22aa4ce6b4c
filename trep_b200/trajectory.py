"""Trajectory files: the reference's MATLAB format, extended to batches.

Same keys and conventions as ``trep.save_trajectory`` / ``trep.load_trajectory``
(trep/system.py:1209-1304): ``time`` [K], ``Q`` [K][nQ], ``p`` [K][nQd], ``v`` [K][nQk],
``u`` [K-1][nu], ``rho`` [K-1][nQk] plus the name indices ``Q_index`` ... ``rho_index`` (cell arrays of
strings), so a file written here loads in the reference (and in MATLAB) and vice versa.  Batches of
rollouts (what ``trepb_step_batch`` / ``trepb_project_batch`` produce) are stored with one leading
axis, ``Q`` [B][K][nQ] etc.; on load, columns are matched to the system's configs / inputs BY NAME
along the last axis exactly as the reference does, missing names are left zero.
"""
from __future__ import annotations

import numpy as np


def _names(system):
    """(configs, dynamic configs, kinematic configs, inputs) names of a SystemDesc, a model-mirror
    System or a live reference trep.System."""
    if hasattr(system, "config_names"):                      # SystemDesc
        cfg = list(system.config_names)
        return cfg, cfg[:system.nd], cfg[system.nd:], list(system.input_names)
    cfg = [c.name for c in system.configs]
    dyn = [c.name for c in getattr(system, "dyn_configs", [c for c in system.configs if not c.kinematic])]
    kin = [c.name for c in getattr(system, "kin_configs", [c for c in system.configs if c.kinematic])]
    return cfg, dyn, kin, [u.name for u in system.inputs]


def save_trajectory(filename, system, t, Q=None, p=None, v=None, u=None, rho=None):
    """Write a trajectory (2-D arrays, the reference's layout) or a batch of trajectories (3-D arrays,
    leading batch axis) to a MATLAB file."""
    import scipy.io
    data = {"time": np.array(t)}
    for key, val in (("Q", Q), ("p", p), ("v", v), ("u", u), ("rho", rho)):
        if val is not None:
            data[key] = np.array(val)
    cfg, dyn, kin, inp = _names(system)
    for key, names in (("Q_index", cfg), ("p_index", dyn), ("v_index", kin), ("u_index", inp), ("rho_index", kin)):
        data[key] = np.array(names, dtype=object)
    scipy.io.savemat(filename, data)


def load_trajectory(filename, system=None):
    """Returns ``(t, Q, p, v, u, rho)`` rearranged to ``system``'s layout (any of them None when absent
    from the file), or - without a system - ``(t, (Q_index, Q), (p_index, p), ...)`` as stored."""
    import scipy.io
    data = scipy.io.loadmat(filename)
    t = data["time"].squeeze()
    raw = {k: data.get(k, None) for k in ("Q", "p", "v", "u", "rho")}
    index = {k: [str(s[0]).strip() if np.size(s) else "" for s in data[k + "_index"].ravel()]
             for k in ("Q", "p", "v", "u", "rho")}
    if system is None:
        return (t,) + tuple((index[k], raw[k]) for k in ("Q", "p", "v", "u", "rho"))
    cfg, dyn, kin, inp = _names(system)
    want = {"Q": cfg, "p": dyn, "v": kin, "u": inp, "rho": kin}
    out = []
    for k in ("Q", "p", "v", "u", "rho"):
        a = raw[k]
        if a is None:
            out.append(None)
            continue
        a = np.asarray(a, dtype=float)
        if len(index[k]) == 0:
            a = a.reshape(a.shape[:-1] + (0,)) if a.ndim else a
        res = np.zeros(a.shape[:-1] + (len(want[k]),))
        for j, name in enumerate(want[k]):
            if name in index[k]:
                res[..., j] = a[..., index[k].index(name)]
        out.append(res)
    return (t,) + tuple(out)
