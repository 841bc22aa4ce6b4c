"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/trepb.h
declares, validates descriptions, generates specialised system types, and refuses to compute
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import golden_util as G
from trep_b200 import build, desc as D, lib, systems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "trepb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trepb_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    syms = header_symbols()
    assert len(syms) >= 25
    raw = lib.raw()
    missing = [s for s in syms if not hasattr(raw, s)]
    assert not missing, missing
    assert set(lib.EXPORTS) == set(syms)
    assert raw.trepb_abi_version() == 2


def test_ctypes_struct_layout_matches_header():
    # field order/size of the argument structs as declared in include/trepb.h (LP64)
    assert C.sizeof(lib.StepArgs) == 8 + 4 + 4 + 3 * 8 + 11 * 8 + 4 + 4 + 2 * 8 + 8
    assert C.sizeof(lib.LinArgs) == 8 + 4 + 4 + 8 + 2 * 8 + 2 * 8 + 6 * 8 + 5 * 8 + 2 * 8 + 12 * 8
    assert C.sizeof(lib.LqrArgs) == 8 + 6 * 4 + 7 * 8
    assert C.sizeof(lib.ProjectArgs) == 8 + 4 + 4 + 3 * 8 + 3 * 8 + 4 + 4 + 5 * 8 + 8
    assert C.sizeof(D.CSysDesc) == 10 * 4 + 17 * 8


@pytest.mark.parametrize("name", G.ALL)
def test_descriptions_validate_and_hash(name):
    d = G.desc(name)
    lib.validate(d)
    h = lib.desc_hash(d)
    assert h != 0
    assert h == lib.desc_hash(D.SystemDesc.from_json(d.to_json())), "hash must survive a JSON round trip"


def test_specialised_registry_and_codegen():
    names = lib.specialized_names()
    for n in build.AOT_SYSTEMS:
        assert n in names
    text, h = build.codegen(G.desc("pend_on_cart1"))
    assert "struct CtSys" in text and "kND = 2" in text and "kNU = 1" in text
    assert h == lib.struct_hash(G.desc("pend_on_cart1")) and ("0x%016x" % h) in text
    # structure is compiled in, numbers are run-time parameters (slots of the kernel's argument block):
    # another cart mass is the SAME specialisation, another description (desc_hash) ...
    assert "par.v[" in text and "kNPAR" in text and "10.0" not in text
    s = systems.pend_on_cart()
    s.world_frame.children[0].set_mass(11.0)
    assert lib.struct_hash(s.describe()) == h
    assert lib.desc_hash(s.describe()) != lib.desc_hash(G.desc("pend_on_cart1"))
    # ... while a change of structure is not: a rotational inertia where there was none, one more input
    s.world_frame.children[0].set_mass(10.0, 0.3, 0.0, 0.0)
    assert lib.struct_hash(s.describe()) != h
    assert lib.struct_hash(G.desc("pend_on_cart2")) != h


def test_cooperative_shapes_and_generated_instantiation():
    """trepb_coop_dims: the link-level shape of a system; the build instantiates the
    compile-time-size cooperative kernels for exactly that shape."""
    assert lib.coop_dims(G.desc("puppet")) == (22, 18, 0, 6, 34, 12, 157, 10, 20, 38, 0)
    assert lib.coop_dims(G.desc("pendulum5")) == (5, 0, 0, 0, 5, 0, 15, 5, 0, 0, 0)
    # LinearSpring / LinearDamper / wrenches: the eleventh entry asks for the run-time tail of the layout
    assert lib.coop_dims(G.desc("dual_pendulums"))[10] == 1 and lib.coop_dims(G.desc("wrench_arm"))[10] == 1
    for name in build.COOP_AOT_SYSTEMS:
        text = open(os.path.join(build.GEN, "coop_%s.cu" % name)).read()
        dims = ", ".join(str(v) for v in lib.coop_dims(G.desc(name)))
        assert "CtDims<%s>" % dims in text
    # tests/host_math_check.cc runs the same instantiation on the CPU
    assert "CtDims<%s>" % ", ".join(str(v) for v in lib.coop_dims(G.desc("puppet"))[:10]) in \
        open(os.path.join(ROOT, "tests", "host_math_check.cc")).read()


def test_invalid_descriptions_are_refused():
    d = G.desc("damped_pendulum")
    bad = D.SystemDesc.from_json(d.to_json())
    bad.pot_kind = np.array([9], np.int32)
    with pytest.raises(lib.TrepbError) as e:
        lib.validate(bad)
    assert "no device implementation" in str(e.value)
    bad = D.SystemDesc.from_json(d.to_json())
    bad.frame_parent = np.array([-1, 2, 1, 2], np.int32)
    with pytest.raises(lib.TrepbError):
        lib.validate(bad)


@pytest.mark.skipif(lib.device_count() > 0, reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_a_device():
    with pytest.raises(lib.TrepbError) as e:
        lib.System(G.desc("damped_pendulum"))
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "trep_b200")
    pat = re.compile(r"^\s*(from|import)\s+.*(oracle|ref_systems|build_ref|hostmath)", re.M)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(root, f)).read()), f
            if f.endswith((".cu", ".cuh", ".h", ".cc")):
                assert "oracle" not in open(os.path.join(root, f)).read(), f


def test_every_batch_entry_point_opens_an_nvtx_range():
    """the reference brackets its hot calls with callgrind markers
    (trep/_trep/midpointvi.c:2674-2738); here every batch entry point of the C ABI opens an NVTX range named
    after itself (trepb_nvtx.h)."""
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "trep_b200", "csrc")
    text = "".join(open(os.path.join(csrc, f)).read() for f in ("trepb_api.cu", "trepb_lqr.cu", "trepb_comm.cu"))
    marked = set(re.findall(r'TREPB_NVTX\("(trepb_[a-z0-9_]+)"\)', text))
    from trep_b200 import lib
    want = {n for n in lib.EXPORTS if n.endswith("_batch") or n.endswith("_batch_dev")}
    assert want and want <= marked, sorted(want - marked)


def test_plugin_adds_a_specialised_kernel_for_a_user_system():
    """A system the library was not built with (the LinearDamper pair of tests/golden/damper_only.npz) gets its
    register-resident kernels from a plug-in: build_plugin generates and compiles the specialisation unit,
    trepb_load_plugin registers it (no GPU needed for either)."""
    d = systems.named_desc("damper_only")
    before = [lib.raw().trepb_specialized_name(i) for i in range(lib.raw().trepb_num_specialized())]
    assert b"damper_only" not in before
    path = build.build_plugin(d, "damper_only")
    assert os.path.exists(path)
    assert lib.load_plugin(path) == 1
    names = [lib.raw().trepb_specialized_name(i) for i in range(lib.raw().trepb_num_specialized())]
    assert b"damper_only" in names
    with pytest.raises(Exception):
        lib.load_plugin(path)            # already loaded: registers nothing
    # kind="coop": compile-time-size cooperative kernels for a shape the library was not built with
    path = build.build_plugin(systems.named_desc("rod"), "rod_coop", kind="coop")
    assert lib.load_plugin(path) == 1
    # ... also for a shape with LinearSprings (their counts stay run-time data: CtDims<..., 1>)
    path = build.build_plugin(systems.named_desc("spring_arms"), "spring_arms_coop", kind="coop", ext=True)
    assert "38, 1>" not in open(os.path.join(build.GEN, "plugin_rod_coop.cu")).read()
    assert ", 1>;" in open(os.path.join(build.GEN, "plugin_spring_arms_coop.cu")).read()
    assert lib.load_plugin(path) == 2       # the plain flavour and the external-slab one
