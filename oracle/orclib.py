"""ctypes binding of oracle/_build/liboracle*.so.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (cfd_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


class Params(C.Structure):  # struct orc::Params
    _fields_ = [(n, C.c_double) for n in (
        "FSAFE", "U_inf", "V_inf", "MACH_inf", "T_inf", "RHO_inf", "P_inf", "C_inf", "FMU", "FGX", "FGY", "QH",
        "FK", "FR", "FCv", "GAMA", "CTE")] + [("XREF", C.c_double * 10), ("YREF", C.c_double * 10)] + [
        (n, C.c_int) for n in ("IRESTART", "MAXITER", "IPRINT", "MOVIE", "ITLOCAL", "MOVING", "NGAS", "use_gcl")]


class BC(C.Structure):  # struct orc_bc
    _fields_ = [
        ("nfixrho", C.c_int), ("ifixrho_node", C.c_void_p), ("rfixrho_value", C.c_void_p),
        ("nfixv", C.c_int), ("ifixv_node", C.c_void_p), ("rfixv_valuex", C.c_void_p), ("rfixv_valuey", C.c_void_p),
        ("nwall", C.c_int), ("wall", C.c_void_p),
        ("nfixt", C.c_int), ("ifixt_node", C.c_void_p), ("rfixt_value", C.c_void_p),
        ("nsets", C.c_int), ("iset_n1", C.c_void_p), ("iset_n2", C.c_void_p), ("iset_elem", C.c_void_p), ("iset_id", C.c_void_p),
        ("nmove", C.c_int), ("i_m", C.c_void_p),
        ("nfix_move", C.c_int), ("ifm", C.c_void_p),
    ]


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    if force or not os.path.exists(so) or not os.path.exists(os.path.join(_HERE, "_build", "liboracle_omp.so")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


_libs = {}


def lib(omp=False):
    key = "omp" if omp else "seq"
    if key in _libs:
        return _libs[key]
    build()
    L = C.CDLL(os.path.join(_HERE, "_build", "liboracle_omp.so" if omp else "liboracle.so"))
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int, _dp, _dp, _ip, C.POINTER(BC)]
    for f in ("orc_destroy", "orc_init"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = None
    L.orc_step.argtypes = [C.c_void_p, C.c_int]
    L.orc_rk_stage.argtypes = [C.c_void_p, C.c_int]
    L.orc_step_part1.argtypes = [C.c_void_p]
    L.orc_step_part1.restype = C.c_double
    L.orc_step_part2.argtypes = [C.c_void_p, C.c_double]
    L.orc_step_part3.argtypes = [C.c_void_p]
    L.orc_geometry.argtypes = [C.c_void_p, C.c_int]
    L.orc_fluid_structure.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.orc_force_visc.argtypes = [C.c_void_p]
    L.orc_residual_norms.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_long)]
    L.orc_scalar.argtypes = [C.c_void_p, C.c_char_p]
    L.orc_scalar.restype = C.c_double
    L.orc_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    for f in ("orc_pow15", "orc_pow05", "orc_powm05"):
        getattr(L, f).argtypes = [C.c_double]
        getattr(L, f).restype = C.c_double
    L.orc_div3_mismatches.argtypes = [C.c_long, C.c_ulonglong]
    L.orc_div3_mismatches.restype = C.c_long
    L.orc_canon_sum.argtypes = [C.c_long, _dp]
    L.orc_canon_sum.restype = C.c_double
    L.orc_vecdot.argtypes = [C.c_int, _dp, _dp]
    L.orc_vecdot.restype = C.c_double
    L.orc_get_esup.argtypes = [_ip, C.c_int, C.c_int, _ip, _ip]
    L.orc_get_psup.argtypes = [_ip, C.c_int, C.c_int, _ip, C.c_int, _ip]
    L.orc_deriv.argtypes = [_dp, _dp, _ip, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    L.orc_masas.argtypes = [_dp, _ip, C.c_int, C.c_int, _dp]
    L.orc_normales.argtypes = [_ip, C.c_int, _dp, _dp, C.c_int, _ip, _dp, _dp]
    L.orc_deltat.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp] + [C.c_double] * 4 + [_dp, _dp]
    L.orc_estab.argtypes = [C.c_int, _ip] + [_dp] * 9 + [C.c_double] * 4 + [_dp] * 4
    L.orc_calcrhs.argtypes = [_dp] * 12 + [_ip, C.c_int, C.c_int] + [C.c_double] * 6
    L.orc_fuente.argtypes = [_dp] * 8 + [_ip, C.c_int]
    L.orc_spmv.argtypes = [_dp, _ip, _ip, _dp, _dp, C.c_int]
    L.orc_bicg.argtypes = [_dp, _ip, _ip, _dp, _dp, _dp, _dp, _ip, C.c_int, C.c_int]
    L.orc_laplace.argtypes = [_ip, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _ip, _ip, _dp, _dp, C.c_int]
    L.orc_gcl_main.argtypes = [_dp] * 9 + [_ip, C.c_int, C.c_int, C.c_double]
    L.orc_smoothing.argtypes = [_dp, _dp, _ip, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_int, C.c_int]
    L.orc_omp_threads.restype = C.c_int
    L.orc_set_omp_threads.argtypes = [C.c_int]
    L.orc_set_omp_threads.restype = C.c_int
    _libs[key] = L
    return L


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


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Full-solver handle around a cfd_b200.deck.LoadedCase."""

    def __init__(self, lc, omp=False, use_gcl=0):
        self.L = lib(omp)
        self.lc = lc
        s = lc.sets
        self._keep = [np.ascontiguousarray(s[:, k]) for k in range(4)] if s.size else [np.zeros(0, np.int32)] * 4
        bc = BC(
            lc.ifixrho_node.size, _vp(lc.ifixrho_node), _vp(lc.rfixrho_value),
            lc.ifixv_node.size, _vp(lc.ifixv_node), _vp(lc.rfixv_valuex), _vp(lc.rfixv_valuey),
            lc.wall.shape[0], _vp(lc.wall),
            lc.ifixt_node.size, _vp(lc.ifixt_node), _vp(lc.rfixt_value),
            s.shape[0], _vp(self._keep[1]), _vp(self._keep[2]), _vp(self._keep[0]), _vp(self._keep[3]),
            lc.i_m.size, _vp(lc.i_m), lc.ifm.size, _vp(lc.ifm),
        )
        self.par = make_params(lc.par, use_gcl)
        self.h = self.L.orc_create(C.byref(self.par), lc.npoin, lc.nelem, lc.X, lc.Y, lc.inpoel, C.byref(bc))
        self.L.orc_init(self.h)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def view(self, name):
        """numpy view (no copy) of an internal array; writes go straight into the oracle state."""
        ptr, n = C.c_void_p(), C.c_long()
        t = self.L.orc_field(self.h, name.encode(), C.byref(ptr), C.byref(n))
        if t < 0:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0, np.float64 if t == 0 else np.int32)
        ct = C.c_double if t == 0 else C.c_int
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n.value,))

    def get(self, name):
        return self.view(name).copy()

    def set(self, name, value):
        v = self.view(name)
        v[:] = np.asarray(value).ravel()

    def scalar(self, name):
        return self.L.orc_scalar(self.h, name.encode())

    def set_scalar(self, name, v):
        self.L.orc_set_scalar(self.h, name.encode(), float(v))

    def step(self, n=1):
        self.L.orc_step(self.h, n)

    def rk_stage(self, irk):
        self.L.orc_rk_stage(self.h, irk)

    def step_part1(self):
        return self.L.orc_step_part1(self.h)

    def step_part2(self, dtmin_global):
        self.L.orc_step_part2(self.h, dtmin_global)

    def step_part3(self):
        self.L.orc_step_part3(self.h)

    def norms(self):
        er, err = np.zeros(4), np.zeros(4)
        self.L.orc_residual_norms(self.h, er, err)
        return er, err

    def fluid_structure(self, dtmin, time):
        self.L.orc_fluid_structure(self.h, float(dtmin), float(time))

    def geometry(self, moving_step=0):
        self.L.orc_geometry(self.h, int(moving_step))

    def step_norms(self):
        """ER, ERR as the time loop evaluated them on its last print step (before U = U1)."""
        return self.get("ER"), self.get("ERR")

    def force_visc(self):
        """FORCE_VISC on the current state -> (F_VX(10), F_VY(10), skin, edge mid x, press/82713.27)"""
        self.L.orc_force_visc(self.h)
        return tuple(self.get(n) for n in ("F_VX", "F_VY", "skin", "skin_x", "skin_p"))
