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


# ------------------------------------------------------------------ struct layout, pinned from C
def _probe(tmp_path):
    """tests/abi/abi_probe.c compiled as C11 against the header: sizeof / offsetof of weno5_params, enum values."""
    import json
    exe = str(tmp_path / "abi_probe")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi", "abi_probe.c"), "-o", exe])
    return json.loads(subprocess.check_output([exe], text=True))


def test_params_layout_matches_ctypes_and_julia(tmp_path):
    """A reordered or retyped field of weno5_params must fail here: the C layout (gcc, from the header itself) is
    compared with the ctypes mirror, with the table committed next to the Julia module, and with the layout that
    Julia computes for the `struct Weno5Params` text in Weno5GPU.jl (isbits struct, C layout rules)."""
    import json
    c = _probe(tmp_path)
    # ctypes mirror
    assert C.sizeof(L.Params) == c["sizeof_weno5_params"] and C.alignment(L.Params) == c["alignof_weno5_params"]
    assert [n for n, _ in L.Params._fields_] == list(c["fields"])
    for name, _ in L.Params._fields_:
        d = getattr(L.Params, name)
        assert (d.offset, d.size) == (c["fields"][name]["offset"], c["fields"][name]["size"]), name
    assert (L.WENO5_OK, L.ERR_ARG, L.ERR_CUDA, L.ERR_NCCL, L.ERR_STATE, L.ERR_NUMERIC) == tuple(
        c["status"][k] for k in ("OK", "ERR_ARG", "ERR_CUDA", "ERR_NCCL", "ERR_STATE", "ERR_NUMERIC"))
    assert (L.BC_OUTFLOW, L.BC_PERIODIC, L.NBND) == (c["bc"]["OUTFLOW"], c["bc"]["PERIODIC"], c["WENO5_NBND"])
    # committed table next to the Julia module
    with open(os.path.join(ROOT, "julia_b200", "julia", "abi_layout.json")) as f:
        assert json.load(f) == c
    # the Julia struct as written
    jl = open(os.path.join(ROOT, "julia_b200", "julia", "Weno5GPU.jl")).read()
    body = re.search(r"struct Weno5Params\n(.*?)\nend", jl, flags=re.S).group(1)
    fields = re.findall(r"(\w+)::(Cint|Cdouble)", body)
    size = {"Cint": 4, "Cdouble": 8}
    off, layout = 0, {}
    for name, ty in fields:
        off = (off + size[ty] - 1) // size[ty] * size[ty]
        layout[name] = {"offset": off, "size": size[ty]}
        off += size[ty]
    assert layout == c["fields"]
    assert (off + 7) // 8 * 8 == c["sizeof_weno5_params"]
    assert "const NBnd = 3" in jl and "BC_OUTFLOW, BC_PERIODIC = Cint(0), Cint(1)" in jl
    # the 3-D extension's struct: C layout vs ctypes vs the Julia text
    assert C.sizeof(L.Params3D) == c["sizeof_weno5_params3d"] and C.alignment(L.Params3D) == c["alignof_weno5_params3d"]
    assert [n for n, _ in L.Params3D._fields_] == list(c["fields3d"])
    for name, _ in L.Params3D._fields_:
        d = getattr(L.Params3D, name)
        assert (d.offset, d.size) == (c["fields3d"][name]["offset"], c["fields3d"][name]["size"]), name
    body = re.search(r"struct Weno5Params3D\n(.*?)\nend", jl, flags=re.S).group(1)
    off, layout = 0, {}
    for name, ty in re.findall(r"(\w+)::(Cint|Cdouble)", body):
        off = (off + size[ty] - 1) // size[ty] * size[ty]
        layout[name] = {"offset": off, "size": size[ty]}
        off += size[ty]
    assert layout == c["fields3d"]
    assert (off + 7) // 8 * 8 == c["sizeof_weno5_params3d"]


def test_pure_c_driver_links_against_the_library(built, tmp_path):
    """tests/abi/c_driver.c -- the ccall sequence of INTEGRATION.md in plain C -- compiles as C11 and links against
    libweno5mhd.so alone (its -m gpu twin in test_reference_pin.py runs it on the reference's shock-cloud problem)."""
    exe = str(tmp_path / "c_driver")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi", "c_driver.c"), "-o", exe, "-L", os.path.dirname(L.LIB_PATH),
                           "-lweno5mhd", "-Wl,-rpath," + os.path.dirname(L.LIB_PATH)])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    # and the 3-D extension's driver
    exe3 = str(tmp_path / "c_driver3d")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi", "c_driver3d.c"), "-o", exe3, "-L", os.path.dirname(L.LIB_PATH),
                           "-lweno5mhd", "-Wl,-rpath," + os.path.dirname(L.LIB_PATH)])
    r = subprocess.run([exe3], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
