"""CPU checks of the drop-in boundary: the shared library builds for sm_100a, loads, exports
exactly the symbols include/weno5mhd.h declares, and fails loudly (no CPU fallback) without a
GPU.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

from julia_b200 import lib as L
from julia_b200 import problems as pr
from util import ROOT

HEADER = os.path.join(ROOT, "include", "weno5mhd.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(weno5_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in weno5mhd.h but not exported"
    assert sorted(L.SIGNATURES) == names, "ctypes table and header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    exported = set(re.findall(r" T (weno5_[a-z0-9_]+)", out))
    assert exported == set(names), exported ^ set(names)


def test_library_has_sm100a_code_and_no_oracle(built):
    out = subprocess.check_output(["cuobjdump", "-lelf", L.LIB_PATH], text=True)
    assert "sm_100a" in out
    deps = subprocess.check_output(["ldd", L.LIB_PATH], text=True)
    assert "oracle" not in deps
    syms = subprocess.check_output(["nm", "-D", L.LIB_PATH], text=True)
    assert "orc_" not in syms        # the product never links the checker


def test_version_and_error_string(built):
    lib = L.load()
    assert lib.weno5_version() == 100
    assert isinstance(lib.weno5_last_error(), bytes)


def _has_gpu():
    return L.load().weno5_device_count() > 0


def test_no_cpu_fallback(built):
    if _has_gpu():
        pytest.skip("a GPU is present")
    from julia_b200.solver import Solver
    g, q, bx, by = pr.orszag_tang(16)
    with pytest.raises(L.Weno5Error) as e:
        Solver(g)
    assert e.value.code == L.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_argument_validation(built):
    lib = L.load()
    h = C.c_void_p()
    bad = L.Params(3, 3, 0.1, 0.1, 5 / 3, 0.8, 1e-6, 0, 0, 1)       # grid too small (the reference ships 16 x 4)
    assert lib.weno5_create(C.byref(bad), 0, 3, 0, C.byref(h)) == L.ERR_ARG
    bad = L.Params(32, 32, 0.1, 0.1, 5 / 3, 0.8, 1e-6, 0, 7, 1)     # unknown BC
    assert lib.weno5_create(C.byref(bad), 0, 32, 0, C.byref(h)) == L.ERR_ARG
    bad = L.Params(32, 32, 0.1, 0.1, 5 / 3, 0.8, 1e-6, 0, 0, 1)     # slab outside the grid
    assert lib.weno5_create(C.byref(bad), 0, 24, 16, C.byref(h)) == L.ERR_ARG
    assert lib.weno5_rk4_step(None, 0.1) == L.ERR_ARG
    assert b"null" in lib.weno5_last_error()
