"""CPU tests of the drop-in boundary: libb200md.so loads, exports every symbol include/b200md.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import pytest
from conftest import ROOT
from openmm_b200 import _lib


def _declared():
    txt = open(os.path.join(ROOT, "include", "b200md.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200md_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.b200md_create(C.byref(h), 0, 16) != 0
    assert b"no CPU fallback" in lib.b200md_last_error(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "openmm_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(base, f)).read()
                assert "oracle" not in txt.replace("the oracle harness", "").replace("oracle harness", "").replace("against the oracle", "").lower() \
                    or f == "systems.py", os.path.join(base, f)


def test_constraint_classification_dry_run_needs_no_device():
    """b200md_check_constraints (what Platform::contextCreated calls before anything is allocated): rigid waters, X-H_n
    clusters and general networks (CCMA) are accepted; a constraint between a massless and a massive particle is not
    (ContextImpl.cpp:86-87), one between two massless particles is ignored (ReferenceConstraints.cpp:60)."""
    import numpy as np
    lib = _lib.load()
    msg = C.create_string_buffer(256)

    def check(mass, cons):
        mass = np.asarray(mass, dtype=np.float64)
        p1 = np.asarray([c[0] for c in cons], dtype=np.int32); p2 = np.asarray([c[1] for c in cons], dtype=np.int32)
        d = np.asarray([c[2] for c in cons], dtype=np.float64)
        D, I = C.POINTER(C.c_double), C.POINTER(C.c_int)
        return lib.b200md_check_constraints(len(mass), mass.ctypes.data_as(D), len(cons), p1.ctypes.data_as(I), p2.ctypes.data_as(I), d.ctypes.data_as(D), msg, 256)

    assert check([16, 1, 1], [(0, 1, 0.0957), (0, 2, 0.0957), (1, 2, 0.1514)]) == 0                  # SETTLE
    assert check([12, 1, 1, 1], [(0, 1, 0.109), (0, 2, 0.109), (0, 3, 0.109)]) == 0                   # X-H3
    assert check([1]*6, [(i, i+1, 1.0) for i in range(5)]) == 0                                       # chain: CCMA
    assert check([12, 1, 1, 1, 1], [(0, k, 0.109) for k in range(1, 5)]) == 0                         # CH4: CCMA
    assert check([0, 1], [(0, 1, 1.5)]) != 0 and b"massless" in msg.value
    assert check([0, 0], [(0, 1, 1.5)]) == 0
    assert check([1, 1], [(0, 5, 1.0)]) != 0
