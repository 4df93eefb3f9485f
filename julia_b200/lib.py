"""
ctypes binding of libweno5mhd.so (include/weno5mhd.h).

This is the Python twin of julia_b200/julia/Weno5GPU.jl: the same C entry points a Julia
host reaches with ``ccall``.  There is no CPU fallback: if the CUDA library is missing, or
no sm_100 GPU is usable, calls raise ``Weno5Error``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# WENO5_LIB: development knob to time an experimental build of the same sources (never a different backend)
LIB_PATH = os.environ.get("WENO5_LIB") or os.path.join(_HERE, "libweno5mhd.so")

WENO5_OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_STATE, ERR_NUMERIC = 0, -1, -2, -3, -4, -5
BC_OUTFLOW, BC_PERIODIC = 0, 1
NBND = 3


class Weno5Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libweno5mhd error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """weno5_params -- the reference's module constants (Weno5_2D.jl:7-29)."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("gamma", C.c_double), ("cfl", C.c_double), ("eps", C.c_double),
                ("bc_x", C.c_int), ("bc_y", C.c_int), ("es_roundtrip", C.c_int)]


class Params3D(C.Structure):
    """weno5_params3d -- the same constants for the 3-D extension."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("dz", C.c_double), ("gamma", C.c_double), ("cfl", C.c_double), ("eps", C.c_double),
                ("bc_x", C.c_int), ("bc_y", C.c_int), ("bc_z", C.c_int), ("es_roundtrip", C.c_int)]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# every symbol include/weno5mhd.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "weno5_last_error": (C.c_char_p, []),
    "weno5_version": (C.c_int, []),
    "weno5_device_count": (C.c_int, []),
    "weno5_create": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "weno5_destroy": (C.c_int, [_vp]),
    "weno5_comm_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
    "weno5_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]),
    "weno5_comm_halo_mode": (C.c_int, [_vp]),
    "weno5_upload": (C.c_int, [_vp, _dp, _dp, _dp]),
    "weno5_download": (C.c_int, [_vp, _dp, _dp, _dp]),
    "weno5_upload_1d": (C.c_int, [_vp, _dp]),
    "weno5_download_1d": (C.c_int, [_vp, _dp]),
    "weno5_boundaries": (C.c_int, [_vp]),
    "weno5_timestep": (C.c_int, [_vp, _dp, _dp, _dp]),
    "weno5_rk4_step": (C.c_int, [_vp, C.c_double]),
    "weno5_advance": (C.c_int, [_vp, C.c_int, C.c_double, _dp, C.POINTER(C.c_int)]),
    "weno5_divb": (C.c_int, [_vp, _dp]),
    "weno5_totals": (C.c_int, [_vp, _dp]),
    "weno5_set_es_switch": (C.c_int, [_vp, C.c_int, C.c_double]),
    "weno5_es_switch_count": (C.c_int, [_vp, C.POINTER(C.c_ulonglong)]),
    "weno5_es_switch_mask": (C.c_int, [_vp, _dp]),
    "weno5_arr2img": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_short), _dp]),
    "weno5_create_3d": (C.c_int, [C.POINTER(Params3D), C.c_int, C.POINTER(_vp)]),
    "weno5_destroy_3d": (C.c_int, [_vp]),
    "weno5_upload_3d": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "weno5_download_3d": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "weno5_timestep_3d": (C.c_int, [_vp, _dp, _dp]),
    "weno5_rk4_step_3d": (C.c_int, [_vp, C.c_double]),
    "weno5_advance_3d": (C.c_int, [_vp, C.c_int, C.c_double, _dp, C.POINTER(C.c_int)]),
    "weno5_divb_3d": (C.c_int, [_vp, _dp]),
    "weno5_totals_3d": (C.c_int, [_vp, _dp]),
    "weno5_classical_rk4_step_3d": (C.c_int, [C.POINTER(Params3D), _dp, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp, _dp]),
    "weno5_launch_count_3d": (C.c_ulonglong, [_vp]),
    "weno5_profile_3d": (C.c_int, [_vp, C.c_int]),
    "weno5_profile_read_3d": (C.c_int, [_vp, _dp, C.POINTER(C.c_ulonglong)]),
    "weno5_classical_rk4_step_2d": (C.c_int, [C.POINTER(Params), _dp, _dp, _dp, C.c_double, _dp, _dp, _dp]),
    "weno5_classical_rk4_step_1d": (C.c_int, [C.POINTER(Params), _dp, C.c_double, _dp]),
    "weno5_step_host": (C.c_int, [_vp, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp]),
    "weno5_pipe_create": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.POINTER(_vp)]),
    "weno5_pipe_step": (C.c_int, [_vp, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp, _dp]),
    "weno5_pipe_overlapped": (C.c_int, [_vp]),
    "weno5_pipe_destroy": (C.c_int, [_vp]),
    "weno5_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_ulonglong]),
    "weno5_host_free": (C.c_int, [_vp]),
    "weno5_launch_count": (C.c_ulonglong, [_vp]),
    "weno5_profile": (C.c_int, [_vp, C.c_int]),
    "weno5_profile_read": (C.c_int, [_vp, _dp, C.POINTER(C.c_ulonglong), C.c_int]),
    "weno5_debug_fetch": (C.c_int, [_vp, C.c_char_p, _dp]),
    "weno5_stream": (_vp, [_vp]),
}

_lib = None


def load():
    """Load libweno5mhd.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Weno5Error(ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                       "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise Weno5Error(rc, load().weno5_last_error().decode(errors="replace"))
    return rc


def dptr(a):
    return a.ctypes.data_as(_dp) if a is not None else _dp()
