"""NSComp2D — the reference's driver (PROGRAM NSComp2D, ns2DComp.ALE.f90:8-389) on the GPU.

Mirrors the reference's call structure: `readInputData`/`loadMeshData` are `deck.read_deck` +
`deck.load`; the constructor does what the program does before its time loop (allocate, RESTART
free-stream branch, NORMALES, DERIV, MASAS, laplace); `step(n)` is n passes of the loop; the
reference subroutines are reachable one by one (`rk_stage`, `geometry`, `fluid_structure`,
`residual_norms`) and in call-site form with host arrays (`calcrhs`, `fuente`, `deltat`,
`estab`, `deriv`, `masas`, `normales`, `laplace`, `bicg`, `spmv`, `vecdot`, `gcl_main`).
All state lives in HBM inside libcfdb200.so; numpy arrays only cross at get/set.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi
from .deck import LoadedCase

_INT_FIELDS = {"inpoel", "esup1", "esup2", "psup1", "psup2", "lap_idx", "lap_rowptr", "ilaux", "n_ipoin"}


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class NSComp2D:
    def __init__(self, lc: LoadedCase, device: int = 0, use_gcl: int = 0, init: bool = True, smooth: bool = False,
                 restart_path: str | None = None):
        self.L = capi.lib()
        self.lc = lc
        if smooth:  # ns2DComp.ALE.f90:63-76: SMOOTH_FIX = I_M + IFM nodes, then the one-time mesh optimiser
            self.smoothing_sweeps = capi.smoothing(lc)
        s = lc.sets
        self._keep = [np.ascontiguousarray(s[:, k]) for k in range(4)] if s.size else [np.zeros(0, np.int32)] * 4
        bc = capi.BC(
            lc.ifixrho_node.size, _vp(lc.ifixrho_node), _vp(lc.rfixrho_value),
            lc.ifixv_node.size, _vp(lc.ifixv_node), _vp(lc.rfixv_valuex), _vp(lc.rfixv_valuey),
            lc.wall.shape[0], _vp(lc.wall),
            lc.ifixt_node.size, _vp(lc.ifixt_node), _vp(lc.rfixt_value),
            s.shape[0], _vp(self._keep[1]), _vp(self._keep[2]), _vp(self._keep[0]), _vp(self._keep[3]),
            lc.i_m.size, _vp(lc.i_m), lc.ifm.size, _vp(lc.ifm),
        )
        self.par = capi.make_params(lc.par, use_gcl)
        self.h = C.c_void_p()
        capi.check(self.L.cfdb_create(C.byref(self.h), C.byref(self.par), lc.npoin, lc.nelem, lc.X, lc.Y, lc.inpoel,
                                      C.byref(bc), device))
        self.npoin, self.nelem = lc.npoin, lc.nelem
        # IRESTART == 1 (ns2DComp.ALE.f90:423-431): the state comes from <name>.RST -- never silently a free-stream start
        self._restart_path = restart_path
        if int(lc.par.get("IRESTART", 0)) == 1 and restart_path is None:
            self.close()
            raise capi.CfdbError("the deck says IRESTART = 1: pass restart_path=<name>.RST (ns2DComp.ALE.f90:423-431)")
        if init:
            self.init()

    def init(self):
        """ns2DComp.ALE.f90:59-136 on the device (cfdb_init); the constructor calls it unless init=False."""
        capi.check(self.L.cfdb_init(self.h))
        if int(self.lc.par.get("IRESTART", 0)) == 1:
            self.restart(self._restart_path)

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.cfdb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU -----------------------------------------------------------------------------
    def attach_partition(self, part):
        """Register the ghost-exchange lists of a cfd_b200.partition.LocalPart (this context holds part.lc)."""
        ranks, sp, si, rp, ri = part.halo_arrays()
        self._halo_keep = (ranks, sp, si, rp, ri)
        capi.check(self.L.cfdb_set_halo(self.h, part.n_owned, ranks.size, ranks, sp, si, rp, ri))
        if getattr(part, "red_aligned", False):   # chunk-aligned ownership: global canonical reductions, bit-identical to one GPU
            capi.check(self.L.cfdb_set_reduction_layout(self.h, int(part.gid0), int(part.npoin_global)))
        if getattr(part, "moving", False):
            self.set_option("ale", 1)   # FUENTE and the mesh-velocity terms also on ranks without body edges of their own
        self.part = part

    def comm_init(self, uid: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(uid, 128)
        capi.check(self.L.cfdb_comm_init(self.h, C.cast(buf, C.c_void_p), rank, nranks))

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().cfdb_nccl_unique_id(C.cast(buf, C.c_void_p)))
        return buf.raw

    def halo_exchange(self, name):
        capi.check(self.L.cfdb_halo_exchange(self.h, name.encode()))

    # ---- resident mode -------------------------------------------------------------------
    def step(self, n=1):
        capi.check(self.L.cfdb_step(self.h, n))

    def sync(self):
        capi.check(self.L.cfdb_sync(self.h))

    def step_streamed(self, ins, outs, norms=None):
        """cfdb_step_streamed: one step whose state comes from / goes back to host arrays (dicts keyed U, T, VEL_X, VEL_Y of
        contiguous float64 numpy arrays, ideally pinned; missing keys are not transferred), pipelined: returns at once."""
        def p(d, k):
            a = d.get(k) if d else None
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        n = norms.ctypes.data_as(C.c_void_p) if norms is not None else None
        capi.check(self.L.cfdb_step_streamed(self.h, p(ins, "U"), p(ins, "T"), p(ins, "VEL_X"), p(ins, "VEL_Y"), p(outs, "U"), p(outs, "T"),
                                             p(outs, "VEL_X"), p(outs, "VEL_Y"), n))

    def streamed_wait(self):
        capi.check(self.L.cfdb_streamed_wait(self.h))

    def rk_stage(self, irk):
        capi.check(self.L.cfdb_rk_stage(self.h, irk))

    def geometry(self, moving_step=0):
        capi.check(self.L.cfdb_geometry(self.h, moving_step))

    def fluid_structure(self, dtmin, time):
        capi.check(self.L.cfdb_fluid_structure(self.h, dtmin, time))

    def norms(self):
        er, err = np.zeros(4), np.zeros(4)
        capi.check(self.L.cfdb_residual_norms(self.h, er, err))
        return er, err

    def force_visc(self):
        """FORCE_VISC (ns2DComp.ALE.f90:819-893) now -> (F_VX(10), F_VY(10), skin, edge mid x, press/82713.27)."""
        capi.check(self.L.cfdb_force_visc(self.h))
        return tuple(self.get(n) for n in ("F_VX", "F_VY", "skin", "skin_x", "skin_p"))

    def printflavia(self, path, it, flags=(1, 1, 1, 1, 1, 1, 1), append=False):
        """PRINTFLAVIA (ns2DComp.ALE.f90:701-817): GiD result blocks; flags = RHO, VEL2, MACH, PRES, TEMP, ENER, POS."""
        capi.check(self.L.cfdb_printflavia(self.h, str(path).encode(), int(it), np.asarray(flags, np.int32), int(append)))

    @staticmethod
    def format_real(kind, v, w=0, d=0):
        """one real as Fortran Ew.d ('E'), Fw.d ('F') or list-directed REAL(8) ('L')"""
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().cfdb_format_real(ord(kind), float(v), int(w), int(d), buf, 128))
        return buf.value.decode()

    @staticmethod
    def cnv_record(it, time, r):
        """one line of <name>.cnv ('(I7, 5E14.6)')"""
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().cfdb_format_cnv(int(it), float(time), np.ascontiguousarray(r, np.float64), buf, 128))
        return buf.value.decode()

    def step_norms(self):
        """ER, ERR as evaluated inside the last print step of cfdb_step (before U = U1)."""
        er, err = np.zeros(4), np.zeros(4)
        capi.check(self.L.cfdb_step_norms(self.h, er, err))
        return er, err

    def get(self, name):
        n = self.L.cfdb_field_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        if name in ("n_ipoin", "n_x", "n_y"):
            n = int(self.scalar("n_m"))
        a = np.zeros(n, np.int32 if name in _INT_FIELDS else np.float64)
        capi.check(self.L.cfdb_get(self.h, name.encode(), _vp(a), n))
        return a

    def get_into(self, name, out):
        """Download straight into a caller-owned (e.g. pinned) contiguous numpy buffer."""
        capi.check(self.L.cfdb_get(self.h, name.encode(), _vp(out), out.size))
        return out

    def set_from(self, name, a):
        """Upload straight from a caller-owned (e.g. pinned) contiguous float64 numpy buffer."""
        capi.check(self.L.cfdb_set(self.h, name.encode(), _vp(a), a.size))

    def set(self, name, value):
        a = np.ascontiguousarray(np.asarray(value, np.float64).ravel())
        capi.check(self.L.cfdb_set(self.h, name.encode(), _vp(a), a.size))

    def scalar(self, name):
        v = C.c_double()
        capi.check(self.L.cfdb_get_scalar(self.h, name.encode(), C.byref(v)))
        return v.value

    def set_scalar(self, name, v):
        capi.check(self.L.cfdb_set_scalar(self.h, name.encode(), float(v)))

    # ---- restart files (PRINTREST / RESTART, ns2DComp.ALE.f90:898-917, :423-431) ---------------------------
    def print_rest(self, path):
        """PRINTREST: <name>.RST with (ITER, TIME) and, per node, (U(1:4), T, GAMM) -- unformatted sequential records,
        byte-identical to the reference's (tests/test_reference_callsites.py).  The reference calls PRINTREST before its
        U = U1 (ns2DComp.ALE.f90:254 vs :277), so its file pairs the state the step started from with the temperature it
        ended with; this method writes the consistent pair (U and T of the same step)."""
        from .deck import write_rst

        write_rst(path, int(self.scalar("ITER")), self.scalar("TIME"), self.get("U"), self.get("T"), self.get("GAMM"))

    def restart(self, path):
        """RESTART with IRESTART == 1: U, T, GAMM come from the file; ITER/TIME are read and dropped and VEL_X/VEL_Y are
        not restored, exactly as the reference does (they are left at zero here, the value fresh ALLOCATE memory has)."""
        from .deck import read_rst

        _, _, U, T, G = read_rst(path, self.npoin)
        self.set("U", U)
        self.set("T", T)
        self.set("GAMM", G)
        z = np.zeros(self.npoin)
        self.set("VEL_X", z)
        self.set("VEL_Y", z)

    def set_option(self, name, value):
        """'use_cuarto' / 'true_rk' (SURVEY.md §8f N1, N2); default 0 = the reference's behaviour.  'ale' = 1: the mesh moves
        although this context's own deck has no body sets (ranks of a multi-GPU run, set by attach_partition)."""
        capi.check(self.L.cfdb_set_option(self.h, name.encode(), int(value)))

    @property
    def stream(self):
        return self.L.cfdb_stream(self.h)

    def profile(self, on=True):
        capi.check(self.L.cfdb_profile_enable(self.h, 1 if on else 0))

    def profile_get(self, kernel):
        ms, n = C.c_double(), C.c_int64()
        capi.check(self.L.cfdb_profile_get(self.h, kernel.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self):
        return self.L.cfdb_launch_count(self.h)

    def convergence(self):
        """sqrt(ER_i/ERR_i), the four columns of <name>.cnv (ns2DComp.ALE.f90:199)."""
        er, err = self.norms()
        return [math.sqrt(a / b) if b != 0 else float("nan") for a, b in zip(er, err)]

    # ---- call-site mode (host arrays, reference argument order) ------------------------------
    def calcrhs(self, rhs, U, theta, T, dNx, dNy, area, shoc, dtl, ts1, ts2, ts3, Cv, lambda_ref, mu_ref, gamma0, T_inf, cte):
        capi.check(self.L.cfdb_calcrhs(self.h, rhs, U, theta, T, dNx, dNy, area, shoc, dtl, ts1, ts2, ts3, self.lc.inpoel,
                                       self.nelem, self.npoin, Cv, lambda_ref, mu_ref, gamma0, T_inf, cte))
        return rhs

    def fuente(self, rhs, U, w_x, w_y, dNx, dNy, area, dtl):
        capi.check(self.L.cfdb_fuente(self.h, rhs, U, w_x, w_y, dNx, dNy, area, dtl, self.lc.inpoel, self.nelem, self.npoin))
        return rhs

    def deltat(self, area, T, vel_x, vel_y, w_x, w_y, FSAFE, FR, GAMA, T_inf):
        dtmin, dt = np.zeros(1), np.zeros(self.nelem)
        capi.check(self.L.cfdb_deltat(self.h, dtmin, dt, self.lc.inpoel, area, T, vel_x, vel_y, w_x, w_y, self.nelem,
                                      self.npoin, FSAFE, FR, GAMA, T_inf))
        return dtmin[0], dt

    def estab(self, U, T, vel_x, vel_y, w_x, w_y, GAMM, dNx, dNy, FR, DTMIN, RHOINF, TINF):
        out = [np.zeros(self.nelem) for _ in range(4)]
        capi.check(self.L.cfdb_estab(self.h, U, T, vel_x, vel_y, w_x, w_y, GAMM, dNx, dNy, self.lc.inpoel, self.nelem,
                                     self.npoin, FR, DTMIN, RHOINF, TINF, *out))
        return out

    def deriv(self, X, Y):
        E = self.nelem
        area, HH, HHX, HHY, dNx, dNy, hmin = (np.zeros(E), np.zeros(E), np.zeros(E), np.zeros(E), np.zeros(3 * E),
                                              np.zeros(3 * E), np.zeros(1))
        capi.check(self.L.cfdb_deriv(self.h, X, Y, self.lc.inpoel, E, self.npoin, area, HH, HHX, HHY, dNx, dNy, hmin))
        return area, HH, HHX, HHY, dNx, dNy, hmin[0]

    def masas(self, area):
        M = np.zeros(self.npoin)
        capi.check(self.L.cfdb_masas(self.h, area, self.lc.inpoel, self.nelem, self.npoin, M))
        return M

    def normales(self, X, Y):
        m = C.c_int32()
        nw = max(1, 2 * self.lc.wall.shape[0])
        ip, nx, ny = np.zeros(nw, np.int32), np.zeros(nw), np.zeros(nw)
        capi.check(self.L.cfdb_normales(self.h, self.lc.wall, self.lc.wall.shape[0], X, Y, self.npoin, C.byref(m), ip, nx, ny))
        return m.value, ip[: m.value], nx[: m.value], ny[: m.value]

    def laplace(self, area, dNx, dNy, X, Y):
        nnz = self.L.cfdb_field_size(self.h, b"lap_sparse")
        sp, dg = np.zeros(nnz), np.zeros(self.npoin)
        capi.check(self.L.cfdb_laplace(self.h, self.lc.inpoel, area, dNx, dNy, X, Y, self.nelem, self.npoin, sp, dg))
        return sp, dg

    def bicg(self, A, idx, rowptr, diag, x, b, x_fix, fix_idx):
        it = C.c_int32()
        capi.check(self.L.cfdb_bicg(self.h, A, idx, rowptr, diag, x, b, x_fix, fix_idx, self.npoin, fix_idx.size, C.byref(it)))
        return it.value

    def spmv(self, A, idx, rowptr, v):
        y = np.zeros(v.size)
        capi.check(self.L.cfdb_spmv(self.h, A, idx, rowptr, v, y, v.size, int(rowptr[v.size])))
        return y

    def vecdot(self, x, y):
        r = C.c_double()
        capi.check(self.L.cfdb_vecdot(self.h, x.size, x, y, C.byref(r)))
        return r.value

    def fixvel(self, ifixv_node, rfixv_valuex, rfixv_valuey, vel_x, vel_y):
        capi.check(self.L.cfdb_fixvel(self.h, ifixv_node.size, ifixv_node, rfixv_valuex, rfixv_valuey, vel_x, vel_y, self.npoin))

    def normalvel(self, n_ipoin, n_x, n_y, vel_x, vel_y, w_x, w_y):
        capi.check(self.L.cfdb_normalvel(self.h, n_ipoin.size, n_ipoin, n_x, n_y, vel_x, vel_y, w_x, w_y, self.npoin))

    def fix(self, FR, GAMM, ifixrho_node, rfixrho_value, ifixt_node, rfixt_value, vel_x, vel_y, rho, T, E):
        capi.check(self.L.cfdb_fix(self.h, FR, GAMM, ifixrho_node.size, ifixrho_node, rfixrho_value, ifixt_node.size, ifixt_node,
                                   rfixt_value, vel_x, vel_y, rho, T, E, self.npoin))

    def rk_callsite(self, DTMIN, NRK, BANDERA, GAMM, dtl, U, U1, RHS, RHS1, RHS2, RHS3, T, P, RHO, E, RMACH, VEL_X, VEL_Y, W_X, W_Y,
                    SHOC, T_SUGN1, T_SUGN2, T_SUGN3):
        """RK(DTMIN, NRK, BANDERA, GAMM, dtl) (subrutinas.f90:645) with the module arrays it touches as host arrays"""
        capi.check(self.L.cfdb_rk(self.h, DTMIN, NRK, BANDERA, GAMM, dtl, U, U1, RHS, RHS1, RHS2, RHS3, T, P, RHO, E, RMACH, VEL_X, VEL_Y,
                                  W_X, W_Y, SHOC, T_SUGN1, T_SUGN2, T_SUGN3, self.nelem, self.npoin))

    def mesh_move(self, dtmin, time, X, Y, X1, Y1, W_X, W_Y, P, xpos, ypos):
        """fluidStructure(dtmin, time, SMOOTH_FIX, x1, y1) (meshMove.f90:28) with host arrays -> (FX, FY, RM)"""
        fx, fy, rm = np.zeros(10), np.zeros(10), np.zeros(10)
        capi.check(self.L.cfdb_mesh_move(self.h, dtmin, time, X, Y, X1, Y1, W_X, W_Y, P, xpos, ypos, fx, fy, rm, self.npoin))
        return fx, fy, rm

    def write_forces(self, path):
        capi.check(self.L.cfdb_write_forces(self.h, str(path).encode()))

    def write_desplazamiento(self, path, time, append=False):
        capi.check(self.L.cfdb_write_desplazamiento(self.h, str(path).encode(), float(time), int(append)))

    def write_skin(self, path):
        capi.check(self.L.cfdb_write_skin(self.h, str(path).encode()))

    def gcl_main(self, M, W_x, W_y, W_x_old, W_y_old, area_old, dNx, dNy, area, dt):
        capi.check(self.L.cfdb_gcl_main(self.h, M, W_x, W_y, W_x_old, W_y_old, area_old, dNx, dNy, area, self.lc.inpoel,
                                        self.nelem, self.npoin, dt))
        return M
