"""ctypes binding of include/cfdb.h (libcfdb200.so).  No fallback: import fails loudly if the
library has not been built (python __graft_entry__.py build, or make -C cfd_b200/csrc)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CFDB_LIB_PATH") or os.path.join(_HERE, "libcfdb200.so")   # override: A/B builds (csrc/Makefile: hints)

from ._abi import SYMBOLS, bind as _bind_abi   # generated from the ABI table (tools/gen_abi.py): every symbol include/cfdb.h declares

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
    _bind_abi(L, Params, BC)
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


def smoothing_colored(lc):
    """the parallel, colour-ordered variant of the mesh optimiser (opt-in: not the reference's node order); updates lc.X, lc.Y"""
    n = C.c_int32()
    check(lib().cfdb_smoothing_colored(lc.X, lc.Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem, C.byref(n)))
    return n.value
