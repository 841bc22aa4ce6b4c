"""trep_b200 - B200-native batched MidpointVI (DEL step, first and second derivatives) behind a
C ABI (include/trepb.h).  The CUDA library is built in-tree by ``python -m trep_b200.build``.

    from trep_b200 import MidpointVI, DSystem, systems     # needs libtrepb.so and a GPU to compute
"""
__all__ = ["MidpointVI", "DSystem", "ConvergenceError", "systems", "model", "desc", "trajectory",
           "save_trajectory", "load_trajectory"]


def __getattr__(name):
    # lazy: importing the package must not require the built library (build.py imports it first)
    if name in ("MidpointVI", "ConvergenceError"):
        from . import midpointvi
        return getattr(midpointvi, name)
    if name == "DSystem":
        from .discopt import DSystem
        return DSystem
    if name in ("save_trajectory", "load_trajectory"):
        from . import trajectory
        return getattr(trajectory, name)
    if name in ("systems", "model", "desc", "lib", "build", "discopt", "midpointvi", "trajectory"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
