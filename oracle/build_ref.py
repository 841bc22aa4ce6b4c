#!/usr/bin/env python3
"""Build the UNMODIFIED-NUMERICS reference (`trep` + its `_trep` C extension) into
``oracle/_ref/`` so it can be imported under Python 3.

TEST INFRASTRUCTURE ONLY.  Nothing under ``trep_b200/`` may import this.

The reference (MurpheyLab/trep 1.0.3) is Python-2-only.  This recipe copies the sources
from where they lie (``/root/reference/trep``) into a scratch directory, applies
*mechanical spelling patches only* (CPython-2 C-API names -> CPython-3 names, Python-2
syntax -> Python-3 syntax), and compiles every C file named in the reference's
``setup.py:136-169`` with ``gcc -O2`` (no ``-march=native``, no ``-ffast-math``,
``-ffp-contract=off`` so that no FMA contraction can change a single rounding).
No numeric line of the reference is touched.  Outputs go only to ``oracle/_ref/``
(git-ignored; it travels to the GPU box with the snapshot like any built artefact).

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

C_SOURCES = [
    # the list in the reference's setup.py:136-169
    "midpointvi.c", "system.c", "math-code.c", "frame.c", "_trep.c", "config.c",
    "potential.c", "force.c", "input.c", "constraint.c", "frametransform.c", "spline.c",
    "tapemeasure.c",
    "constraints/distance.c", "constraints/plane.c", "constraints/point.c",
    "potentials/gravity.c", "potentials/linearspring.c", "potentials/configspring.c",
    "potentials/nonlinear_config_spring.c",
    "forces/damping.c", "forces/lineardamper.c", "forces/configforce.c",
    "forces/bodywrench.c", "forces/hybridwrench.c", "forces/spatialwrench.c",
    "forces/pistonexample.c",
]

PY_PACKAGES = ["", "constraints", "potentials", "forces", "discopt", "puppets"]


# ----------------------------------------------------------------------------------------
# C-API spelling patches (CPython 2 -> 3).  None touches arithmetic.
# ----------------------------------------------------------------------------------------
def patch_c(text, name):
    text = re.sub(r"PyObject_HEAD_INIT\(NULL\)\s*\n\s*0,[^\n]*\n",
                  "PyVarObject_HEAD_INIT(NULL, 0)\n", text)
    text = re.sub(r"\(\(PyObject\*\)(\w+)\)->ob_type", r"Py_TYPE(\1)", text)
    text = re.sub(r"(\w+)->ob_type", r"Py_TYPE(\1)", text)
    text = text.replace("PyExc_StandardError", "PyExc_Exception")
    text = text.replace("PyInt_FromLong", "PyLong_FromLong")
    text = text.replace("PyInt_AsLong", "PyLong_AsLong")
    text = text.replace("PyString_FromString", "PyUnicode_FromString")
    text = text.replace("PyCObject_FromVoidPtr((void *)&trep_API_def, NULL)",
                        'PyCapsule_New((void *)&trep_API_def, "trep._C_API", NULL)')
    text = text.replace("PyCObject_Check", "PyCapsule_CheckExact")
    text = text.replace("PyCObject_AsVoidPtr(trep_api_object)",
                        'PyCapsule_GetPointer(trep_api_object, "trep._C_API")')
    if name == "_trep.c":
        text = text.replace("PyMODINIT_FUNC init_trep(void)",
                            "static struct PyModuleDef trep_moduledef = {\n"
                            "    PyModuleDef_HEAD_INIT, \"_trep\", \"trep C core\", -1, CTrepMethods\n};\n"
                            "PyMODINIT_FUNC PyInit__trep(void)")
        text = re.sub(r"m = Py_InitModule3\([^;]*;", "m = PyModule_Create(&trep_moduledef);", text)
        # `return;` -> `return NULL;` inside the init function, and return the module at the end.
        head, sep, tail = text.partition("PyMODINIT_FUNC PyInit__trep(void)")
        tail = re.sub(r"\breturn;", "return NULL;", tail)
        tail = tail.replace("import_array()", "import_array();")
        idx = tail.rindex("}")
        tail = tail[:idx] + "    return m;\n" + tail[idx:]
        text = head + sep + tail
        text = text.replace("#define PyMODINIT_FUNC void", "#define PyMODINIT_FUNC PyObject*")
    return text


# ----------------------------------------------------------------------------------------
# Python 2 -> 3 syntax patches.
# ----------------------------------------------------------------------------------------
def _fix_print(line):
    m = re.match(r"^(\s*)print\s+(.*)$", line)
    if not m:
        m2 = re.match(r"^(\s*)print\s*$", line)
        if m2:
            return m2.group(1) + "print()"
        return line
    indent, rest = m.groups()
    if rest.startswith("("):
        return line
    return "%sprint(%s)" % (indent, rest)


def patch_py(text, relpath, siblings):
    # multi-line print """ ... """ (midpointvi.py:30-34)
    text = re.sub(r'print """(.*?)"""', r'print("""\1""")', text, flags=re.S)
    lines = [_fix_print(l) for l in text.split("\n")]
    text = "\n".join(lines)
    text = text.replace(".iteritems()", ".items()").replace(".itervalues()", ".values()")
    text = re.sub(r"\bxrange\b", "range", text)
    text = re.sub(r"exec (\w+) in (\w+)", r"exec(\1, \2)", text)
    text = re.sub(r"except (\w+), (\w+):", r"except \1 as \2:", text)
    text = re.sub(r"\bnp\.int\b", "int", text)
    text = re.sub(r"\bnp\.float\b", "float", text)
    text = re.sub(r"\bnp\.object\b", "object", text)
    text = text.replace("inspect.getargspec", "inspect.getfullargspec")
    text = text.replace("spec.keywords", "spec.varkw")
    text = re.sub(r"\bStandardError\b", "Exception", text)
    # implicit relative imports
    def fix_from(m):
        mod = m.group(2)
        if mod.split(".")[0] in siblings:
            return "%sfrom .%s import" % (m.group(1), mod)
        return m.group(0)
    text = re.sub(r"^(\s*)from ([\w\.]+) import", fix_from, text, flags=re.M)
    def fix_import(m):
        mod = m.group(2)
        if mod in siblings:
            return "%sfrom . import %s" % (m.group(1), mod)
        return m.group(0)
    text = re.sub(r"^(\s*)import (\w+)\s*$", fix_import, text, flags=re.M)
    if relpath == "puppets/puppets.py":
        # drop the OpenGL / visual half (UI; outside the hot path)
        text = re.sub(r"^from OpenGL.*$", "", text, flags=re.M)
        text = re.sub(r"^from trep\.visual import \*$", "", text, flags=re.M)
        cut = text.find("class PuppetVisual")
        if cut >= 0:
            text = text[:cut]
    if relpath == "forces/__init__.py":
        pass
    return text


LINEARDAMPER_BUG = "dx_dqdq2 = TapeMeasure_length_dq(self->path, q2);"
LINEARDAMPER_FIX = "dx_dqdq2 = TapeMeasure_length_dqdq(self->path, q, q2);"


def build(reference, force=False, out=None, fix_lineardamper=False):
    """`out` / `fix_lineardamper`: a SECOND build outside the repo (oracle/gen_golden_r2.py uses a temp dir) with
    the one-line typo of forces/lineardamper.c:99 corrected (f_ddqdq reads length_dq(q2) where the product rule
    - and its own f_dqdq at :79 - needs length_dqdq(q, q2)).  It exists only to show that this library's
    LinearDamper second derivatives equal the reference's once that line is right; oracle/_ref itself is never
    built with it."""
    OUT = out or globals()["OUT"]
    src_pkg = os.path.join(reference, "trep")
    so_path = os.path.join(OUT, "trep", "_trep" + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(so_path) and not force:
        return OUT
    if not os.path.isdir(src_pkg):
        raise SystemExit("reference sources not found at %s (the GPU box uses the prebuilt "
                         "oracle/_ref)" % src_pkg)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(os.path.join(OUT, "trep"))

    # ---- python layer ------------------------------------------------------------------
    for pkg in PY_PACKAGES:
        sdir = os.path.join(src_pkg, pkg)
        ddir = os.path.join(OUT, "trep", pkg)
        os.makedirs(ddir, exist_ok=True)
        names = [f for f in os.listdir(sdir) if f.endswith(".py")]
        siblings = {f[:-3] for f in names} | {"_trep", "__version__"}
        siblings |= {d for d in os.listdir(sdir)
                     if os.path.isdir(os.path.join(sdir, d)) and d in PY_PACKAGES}
        if pkg != "":
            siblings.discard("_trep")
        for f in names:
            rel = (pkg + "/" + f) if pkg else f
            with open(os.path.join(sdir, f)) as fh:
                text = fh.read()
            with open(os.path.join(ddir, f), "w") as fh:
                fh.write(patch_py(text, rel, siblings))
    with open(os.path.join(OUT, "trep", "__version__.py"), "w") as fh:
        fh.write("__version__ = 'v1.0.3-oracle'\n")

    # ---- C extension -------------------------------------------------------------------
    import numpy
    scratch = tempfile.mkdtemp(prefix="trep_ref_build_")
    try:
        csrc = os.path.join(scratch, "_trep")
        shutil.copytree(os.path.join(src_pkg, "_trep"), csrc)
        for root, _, files in os.walk(csrc):
            for f in files:
                if f.endswith((".c", ".h")):
                    p = os.path.join(root, f)
                    with open(p) as fh:
                        text = fh.read()
                    if fix_lineardamper and f == "lineardamper.c":
                        assert text.count(LINEARDAMPER_BUG) == 1
                        text = text.replace(LINEARDAMPER_BUG, LINEARDAMPER_FIX)
                    with open(p, "w") as fh:
                        fh.write(patch_c(text, f))
        incs = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), "-I" + csrc]
        cflags = ["-O2", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-w",
                  "-DNPY_NO_DEPRECATED_API=0"]
        objs = []
        for s in C_SOURCES:
            o = os.path.join(scratch, s.replace("/", "_") + ".o")
            cmd = ["gcc"] + cflags + incs + ["-c", os.path.join(csrc, s), "-o", o]
            subprocess.check_call(cmd)
            objs.append(o)
        subprocess.check_call(["gcc", "-shared", "-o", so_path] + objs + ["-lpthread", "-lm"])
        # headers a C harness may compile against (the reference installs them too, setup.py:229-232)
        hdir = os.path.join(OUT, "trep", "_trep")
        os.makedirs(hdir, exist_ok=True)
        for h in ("trep.h", "c_api.h"):
            shutil.copy(os.path.join(csrc, h), os.path.join(hdir, h))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return OUT


def import_ref():
    """Import the built reference package (``oracle/_ref/trep``); returns the module."""
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import trep  # noqa
    return trep


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    out = build(a.reference, a.force)
    t = import_ref()
    print("built reference into", out, "->", t.__file__)
