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
