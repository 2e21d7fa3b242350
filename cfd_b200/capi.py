"""ctypes binding of include/cfdb.h (libcfdb200.so).  No fallback: import fails loudly if the
library has not been built (python __graft_entry__.py build, or make -C cfd_b200/csrc)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CFDB_LIB_PATH") or os.path.join(_HERE, "libcfdb200.so")   # override: A/B builds (csrc/Makefile: hints)

_dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")

# every symbol include/cfdb.h declares (tests/test_abi.py checks the library exports each one)
SYMBOLS = [
    "cfdb_last_error", "cfdb_device_count", "cfdb_create", "cfdb_destroy", "cfdb_init", "cfdb_step", "cfdb_sync",
    "cfdb_rk_stage", "cfdb_geometry", "cfdb_fluid_structure", "cfdb_residual_norms", "cfdb_step_norms", "cfdb_force_visc", "cfdb_printflavia", "cfdb_format_cnv", "cfdb_format_real", "cfdb_get", "cfdb_set",
    "cfdb_field_size", "cfdb_get_scalar", "cfdb_set_scalar", "cfdb_set_option", "cfdb_stream", "cfdb_profile_enable", "cfdb_profile_get",
    "cfdb_launch_count", "cfdb_nccl_unique_id", "cfdb_comm_init", "cfdb_set_halo", "cfdb_halo_exchange", "cfdb_set_reduction_layout", "cfdb_calcrhs", "cfdb_fuente", "cfdb_deltat", "cfdb_estab", "cfdb_deriv", "cfdb_masas",
    "cfdb_normales", "cfdb_laplace", "cfdb_bicg", "cfdb_spmv", "cfdb_vecdot", "cfdb_gcl_main", "cfdb_smoothing", "cfdb_selftest", "cfdb_get_esup",
    "cfdb_get_psup",
]


class Params(C.Structure):  # struct cfdb_params
    _fields_ = [(n, C.c_double) for n in (
        "FSAFE", "U_inf", "V_inf", "MACH_inf", "T_inf", "RHO_inf", "P_inf", "C_inf", "FMU", "FGX", "FGY", "QH",
        "FK", "FR", "FCv", "GAMA", "CTE")] + [("XREF", C.c_double * 10), ("YREF", C.c_double * 10)] + [
        (n, C.c_int32) for n in ("IRESTART", "MAXITER", "IPRINT", "MOVIE", "ITLOCAL", "MOVING", "NGAS", "use_gcl")]


class BC(C.Structure):  # struct cfdb_bc
    _fields_ = [
        ("nfixrho", C.c_int32), ("ifixrho_node", C.c_void_p), ("rfixrho_value", C.c_void_p),
        ("nfixv", C.c_int32), ("ifixv_node", C.c_void_p), ("rfixv_valuex", C.c_void_p), ("rfixv_valuey", C.c_void_p),
        ("nwall", C.c_int32), ("wall", C.c_void_p),
        ("nfixt", C.c_int32), ("ifixt_node", C.c_void_p), ("rfixt_value", C.c_void_p),
        ("nsets", C.c_int32), ("iset_n1", C.c_void_p), ("iset_n2", C.c_void_p), ("iset_elem", C.c_void_p), ("iset_id", C.c_void_p),
        ("nmove", C.c_int32), ("i_m", C.c_void_p),
        ("nfix_move", C.c_int32), ("ifm", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built and cfd_b200 has no CPU fallback "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`)")
    L = C.CDLL(LIB_PATH)
    d, i32, i64, vp, cp = C.c_double, C.c_int32, C.c_int64, C.c_void_p, C.c_char_p
    L.cfdb_last_error.restype = cp
    L.cfdb_device_count.restype = C.c_int
    L.cfdb_create.argtypes = [C.POINTER(vp), C.POINTER(Params), i32, i32, _dp, _dp, _ip, C.POINTER(BC), C.c_int]
    L.cfdb_destroy.argtypes = [vp]
    L.cfdb_destroy.restype = None
    L.cfdb_init.argtypes = [vp]
    L.cfdb_step.argtypes = [vp, i32]
    L.cfdb_sync.argtypes = [vp]
    L.cfdb_rk_stage.argtypes = [vp, i32]
    L.cfdb_geometry.argtypes = [vp, i32]
    L.cfdb_fluid_structure.argtypes = [vp, d, d]
    L.cfdb_residual_norms.argtypes = [vp, _dp, _dp]
    L.cfdb_step_norms.argtypes = [vp, _dp, _dp]
    L.cfdb_force_visc.argtypes = [vp]
    L.cfdb_printflavia.argtypes = [vp, cp, i32, _ip, i32]
    L.cfdb_format_cnv.argtypes = [i32, d, _dp, cp, i32]
    L.cfdb_format_real.argtypes = [i32, d, i32, i32, cp, i32]
    L.cfdb_get.argtypes = [vp, cp, vp, i64]
    L.cfdb_set.argtypes = [vp, cp, vp, i64]
    L.cfdb_field_size.argtypes = [vp, cp]
    L.cfdb_field_size.restype = i64
    L.cfdb_get_scalar.argtypes = [vp, cp, C.POINTER(d)]
    L.cfdb_set_scalar.argtypes = [vp, cp, d]
    L.cfdb_set_option.argtypes = [vp, cp, i32]
    L.cfdb_stream.argtypes = [vp]
    L.cfdb_stream.restype = vp
    L.cfdb_profile_enable.argtypes = [vp, i32]
    L.cfdb_profile_get.argtypes = [vp, cp, C.POINTER(d), C.POINTER(i64)]
    L.cfdb_launch_count.argtypes = [vp]
    L.cfdb_launch_count.restype = i64
    L.cfdb_nccl_unique_id.argtypes = [vp]
    L.cfdb_comm_init.argtypes = [vp, vp, i32, i32]
    L.cfdb_set_halo.argtypes = [vp, i32, i32, _ip, _ip, _ip, _ip, _ip]
    L.cfdb_halo_exchange.argtypes = [vp, cp]
    L.cfdb_set_reduction_layout.argtypes = [vp, i64, i64]
    L.cfdb_calcrhs.argtypes = [vp] + [_dp] * 12 + [_ip, i32, i32] + [d] * 6
    L.cfdb_fuente.argtypes = [vp] + [_dp] * 8 + [_ip, i32, i32]
    L.cfdb_deltat.argtypes = [vp, _dp, _dp, _ip] + [_dp] * 6 + [i32, i32] + [d] * 4
    L.cfdb_estab.argtypes = [vp] + [_dp] * 9 + [_ip, i32, i32] + [d] * 4 + [_dp] * 4
    L.cfdb_deriv.argtypes = [vp, _dp, _dp, _ip, i32, i32] + [_dp] * 7
    L.cfdb_masas.argtypes = [vp, _dp, _ip, i32, i32, _dp]
    L.cfdb_normales.argtypes = [vp, _ip, i32, _dp, _dp, i32, C.POINTER(i32), _ip, _dp, _dp]
    L.cfdb_laplace.argtypes = [vp, _ip, _dp, _dp, _dp, _dp, _dp, i32, i32, _dp, _dp]
    L.cfdb_bicg.argtypes = [vp, _dp, _ip, _ip, _dp, _dp, _dp, _dp, _ip, i32, i32, C.POINTER(i32)]
    L.cfdb_spmv.argtypes = [vp, _dp, _ip, _ip, _dp, _dp, i32, i32]
    L.cfdb_vecdot.argtypes = [vp, i32, _dp, _dp, C.POINTER(d)]
    L.cfdb_gcl_main.argtypes = [vp] + [_dp] * 9 + [_ip, i32, i32, d]
    L.cfdb_smoothing.argtypes = [_dp, _dp, _ip, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), i32, i32, C.POINTER(i32)]
    L.cfdb_selftest.argtypes = [vp, i32, i64, C.c_uint64, C.POINTER(i64)]
    L.cfdb_get_esup.argtypes = [_ip, i32, i32, _ip, _ip]
    L.cfdb_get_psup.argtypes = [_ip, i32, i32, _ip, i32, _ip, C.POINTER(i32)]
    _lib = L
    return L


class CfdbError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise CfdbError(lib().cfdb_last_error().decode())


def make_params(par: dict, use_gcl=0) -> Params:
    p = Params()
    for n, _ in Params._fields_:
        if n in ("XREF", "YREF"):
            for i in range(10):
                getattr(p, n)[i] = par[n][i]
        elif n == "use_gcl":
            p.use_gcl = use_gcl
        else:
            setattr(p, n, par[n])
    return p


def get_esup(inpoel, npoin):
    nelem = inpoel.shape[0]
    e1, e2 = np.zeros(3 * nelem, np.int32), np.zeros(npoin + 1, np.int32)
    check(lib().cfdb_get_esup(inpoel, nelem, npoin, e1, e2))
    return e1, e2


def get_psup(inpoel, npoin):
    nelem = inpoel.shape[0]
    cap = 6 * nelem + 16
    p1, p2, cnt = np.zeros(cap, np.int32), np.zeros(npoin + 1, np.int32), C.c_int32()
    check(lib().cfdb_get_psup(inpoel, nelem, npoin, p1, cap, p2, C.byref(cnt)))
    return p1[: cnt.value].copy(), p2


def smoothing(lc):
    """smoothing_mod::smoothing on a LoadedCase (updates lc.X, lc.Y in place); returns the number of sweeps."""
    n = C.c_int32()
    check(lib().cfdb_smoothing(lc.X, lc.Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem, C.byref(n)))
    return n.value
