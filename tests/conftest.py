import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Compile the product library (nvcc cross-compiles without a GPU) and the oracle."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def hostmath():
    """Host build of the product's per-face arithmetic (julia_b200/csrc/weno5_math.cuh)."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "hostmath")
    so, src = os.path.join(d, "libhostmath.so"), os.path.join(d, "hostmath.cpp")
    hdr = os.path.join(ROOT, "julia_b200", "csrc", "weno5_math.cuh")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-mfma", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.hm_sweep.argtypes = [C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp]
    lib.hm_sweep_S.argtypes = lib.hm_sweep.argtypes
    for f in (lib.hm_e2s, lib.hm_s2e):
        f.restype = C.c_double
        f.argtypes = [C.c_double] * 9
    return lib


@pytest.fixture(scope="session")
def host3d():
    """Host build of the product's 3-D kernel bodies and stage sequence (julia_b200/csrc/weno5_3d_body.cuh)."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "hostmath")
    so, src = os.path.join(d, "libhost3d.so"), os.path.join(d, "host3d.cpp")
    hdrs = [os.path.join(ROOT, "julia_b200", "csrc", h) for h in ("weno5_math.cuh", "weno5_3d_body.cuh")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-mfma", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    lib = C.CDLL(so)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.h3_rk4_step.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_double, C.c_double, ip, C.c_int, C.c_int, dp, dp, dp, dp,
                                C.c_double, dp, dp, dp, dp]
    lib.h3_march_segments.argtypes = [C.c_int] * 5
    lib.h3_xwarp_plan.argtypes = [C.c_int] * 5 + [ip, ip]
    lib.h3_timestep.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_double, C.c_double, dp, dp, dp]
    lib.h3_divb_center2face.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
    return lib
