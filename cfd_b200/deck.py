"""Input decks of the reference (the `dataLoader` boundary) — host-side mirror.

The reference reads three list-directed text files (dataLoader.f90:17-65 `readInputData`,
dataLoader.f90:95-275 `loadMeshData`):

  EULER.DAT        line 1 = case name
  <name>-1.dat     scalars (layout in `write_deck`)
  <name>.dat       mesh + boundary-condition lists

`RawCase` holds what the files hold; `load()` applies exactly the post-processing the Fortran
loader applies (scaling by free-stream values, single-precision TWALL — SURVEY.md F11,
NFIXV += NFIXVI, NFIXT += no-slip nodes, ilaux = [I_M; IFM]) and returns the arrays as
`MeshData`/`InputData` hold them after loading.  Indices stay 1-based int32, `inpoel` is
(nelem,3) C-order == Fortran (3,nelem).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

I32 = np.int32
F64 = np.float64


def _i(a, cols=None):
    a = np.asarray(a, dtype=I32)
    if cols is not None:
        a = a.reshape(-1, cols)
    return np.ascontiguousarray(a)


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=F64))


@dataclass
class RawCase:
    """File-level content of one case (values as written in the decks, before scaling)."""

    name: str
    X: np.ndarray
    Y: np.ndarray
    inpoel: np.ndarray  # (nelem,3) 1-based
    # <name>-1.dat scalars
    IRESTART: int = 0
    MAXITER: int = 1000
    IPRINT: int = 1000
    MOVIE: int = 0
    ITLOCAL: int = 0
    FSAFE: float = 0.3
    U_inf: float = 0.0
    V_inf: float = 0.0
    MACH_inf: float = 0.5
    T_inf: float = 288.0
    RHO_inf: float = 1.225
    P_inf: float = 0.0
    FMU: float = 0.0
    FGX: float = 0.0
    FGY: float = 0.0
    QH: float = 0.0
    FK: float = 0.0
    FR: float = 287.0
    FCv: float = 717.5
    GAMA: float = 1.4
    NGAS: int = 0
    CTE: float = 1.0  # file value; the code uses 1/CTE (dataLoader.f90:59)
    MOVING: int = 0
    XREF1: float = 0.0
    YREF1: float = 0.0
    # <name>.dat lists
    fixrho: tuple = (np.zeros(0, I32), np.zeros(0))  # (node, factor)  factor<0 -> 1.225
    fixvi: tuple = (np.zeros(0, I32), np.zeros(0), np.zeros(0))  # inflow (node, fx, fy) x U_inf, V_inf
    fixv: np.ndarray = field(default_factory=lambda: np.zeros(0, I32))  # no-slip nodes
    wall: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), I32))  # slip-wall edges
    fixt: tuple = (np.zeros(0, I32), np.zeros(0))  # (node, factor x T_inf)
    sets: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), I32))  # (elem, n1, n2, set id)
    master: np.ndarray = field(default_factory=lambda: np.zeros(0, I32))
    slave: np.ndarray = field(default_factory=lambda: np.zeros(0, I32))
    ifm: np.ndarray = field(default_factory=lambda: np.zeros(0, I32))  # nodes with fixed (zero) mesh motion
    i_m: np.ndarray = field(default_factory=lambda: np.zeros(0, I32))  # nodes moving with the body

    @property
    def npoin(self):
        return int(self.X.shape[0])

    @property
    def nelem(self):
        return int(self.inpoel.shape[0])


@dataclass
class LoadedCase:
    """State of InputData + MeshData right after readInputData/loadMeshData."""

    name: str
    par: dict  # InputData scalars (CTE inverted, free-stream completed)
    X: np.ndarray
    Y: np.ndarray
    inpoel: np.ndarray
    ifixrho_node: np.ndarray
    rfixrho_value: np.ndarray
    ifixv_node: np.ndarray
    rfixv_valuex: np.ndarray
    rfixv_valuey: np.ndarray
    wall: np.ndarray
    ifixt_node: np.ndarray
    rfixt_value: np.ndarray
    sets: np.ndarray  # (nsets,4): elem, n1, n2, id  (file order)
    ifm: np.ndarray
    i_m: np.ndarray
    ilaux: np.ndarray
    smooth_fix: np.ndarray  # uint8 mask, ns2DComp.ALE.f90:63-73

    @property
    def npoin(self):
        return int(self.X.shape[0])

    @property
    def nelem(self):
        return int(self.inpoel.shape[0])


def load(raw: RawCase) -> LoadedCase:
    """dataLoader.f90:57-64 and :121-268 applied to in-memory deck content."""
    p = dict(
        IRESTART=raw.IRESTART, MAXITER=raw.MAXITER, IPRINT=raw.IPRINT, MOVIE=raw.MOVIE, ITLOCAL=raw.ITLOCAL,
        FSAFE=raw.FSAFE, U_inf=raw.U_inf, V_inf=raw.V_inf, MACH_inf=raw.MACH_inf, T_inf=raw.T_inf,
        RHO_inf=raw.RHO_inf, P_inf=raw.P_inf, FMU=raw.FMU, FGX=raw.FGX, FGY=raw.FGY, QH=raw.QH, FK=raw.FK,
        FR=raw.FR, FCv=raw.FCv, GAMA=raw.GAMA, NGAS=raw.NGAS, MOVING=raw.MOVING,
    )
    p["CTE"] = 1.0 / raw.CTE
    if p["T_inf"] == 0.0:
        p["T_inf"] = p["P_inf"] / (p["FR"] * p["RHO_inf"])
    if p["P_inf"] == 0.0:
        p["P_inf"] = p["RHO_inf"] * p["FR"] * p["T_inf"]
    if p["RHO_inf"] == 0.0:
        p["RHO_inf"] = p["P_inf"] / (p["FR"] * p["T_inf"])
    p["C_inf"] = math.sqrt(p["GAMA"] * p["FR"] * p["T_inf"])
    if math.sqrt(p["U_inf"] ** 2 + p["V_inf"] ** 2) == 0.0:
        p["U_inf"] = p["C_inf"] * p["MACH_inf"]
    xref = [0.0] * 10
    yref = [0.0] * 10
    xref[0], yref[0] = raw.XREF1, raw.YREF1
    p["XREF"], p["YREF"] = xref, yref

    rho_n, rho_f = _i(raw.fixrho[0]), _f(raw.fixrho[1])
    rho_v = np.where(rho_f < 0, 1.225, rho_f * p["RHO_inf"])
    vi_n, vi_x, vi_y = _i(raw.fixvi[0]), _f(raw.fixvi[1]) * p["U_inf"], _f(raw.fixvi[2]) * p["V_inf"]
    ns = _i(raw.fixv)
    # TWALL is implicitly typed single precision (dataLoader.f90:161)
    twall = float(np.float32(p["T_inf"] * (1.0 + (p["GAMA"] - 1) / 2.0 * p["MACH_inf"] * p["MACH_inf"])))
    fixv_n = np.concatenate([vi_n, ns]).astype(I32)
    fixv_x = np.concatenate([vi_x, np.zeros(ns.size)])
    fixv_y = np.concatenate([vi_y, np.zeros(ns.size)])
    t_n, t_v = _i(raw.fixt[0]), _f(raw.fixt[1]) * p["T_inf"]
    fixt_n = np.concatenate([ns, t_n]).astype(I32)
    fixt_v = np.concatenate([np.full(ns.size, twall), t_v])
    if raw.master.size != raw.slave.size:
        raise ValueError("ERROR NODOS MASTER DISTINTO NODOS SLAVE")  # dataLoader.f90:223-226
    i_m, ifm = _i(raw.i_m), _i(raw.ifm)
    ilaux = np.concatenate([i_m, ifm]).astype(I32)
    fix = np.zeros(raw.npoin, np.uint8)
    fix[i_m - 1] = 1
    fix[ifm - 1] = 1
    return LoadedCase(
        name=raw.name, par=p, X=_f(raw.X).copy(), Y=_f(raw.Y).copy(), inpoel=_i(raw.inpoel, 3),
        ifixrho_node=rho_n, rfixrho_value=_f(rho_v), ifixv_node=fixv_n, rfixv_valuex=_f(fixv_x),
        rfixv_valuey=_f(fixv_y), wall=_i(raw.wall, 2), ifixt_node=fixt_n, rfixt_value=_f(fixt_v),
        sets=_i(raw.sets, 4), ifm=ifm, i_m=i_m, ilaux=ilaux, smooth_fix=fix,
    )


# ---------------------------------------------------------------------------------------------
# text formats

def _r(x):
    return repr(float(x))


MESH_MAGIC = b"CFDBMSH1"


def write_mesh_binary(raw: RawCase, path: str) -> None:
    """<name>.cfdbmesh: the content of <name>.dat as little-endian raw arrays, in the order the text file lists them
    (a text deck of a 16 M-triangle mesh is > 1 GB and takes minutes to parse).  Layout: magic, 12 int32
    counts (npoin nelem nfixrho nfixvi nfixv nwall nfixt nsets nmaster nslave nfix_move nmove), X, Y (f64), inpoel (i32,
    3 per element), then per list: node ids (i32) followed by its value columns (f64)."""
    cnt = np.array([raw.npoin, raw.nelem, len(raw.fixrho[0]), len(raw.fixvi[0]), len(raw.fixv), len(raw.wall), len(raw.fixt[0]),
                    len(raw.sets), len(raw.master), len(raw.slave), len(raw.ifm), len(raw.i_m)], I32)
    with open(path, "wb") as f:
        f.write(MESH_MAGIC)
        f.write(cnt.tobytes())
        for a, t in ((raw.X, F64), (raw.Y, F64), (raw.inpoel, I32), (raw.fixrho[0], I32), (raw.fixrho[1], F64),
                     (raw.fixvi[0], I32), (raw.fixvi[1], F64), (raw.fixvi[2], F64), (raw.fixv, I32), (raw.wall, I32),
                     (raw.fixt[0], I32), (raw.fixt[1], F64), (raw.sets, I32), (raw.master, I32), (raw.slave, I32),
                     (raw.ifm, I32), (raw.i_m, I32)):
            f.write(np.ascontiguousarray(np.asarray(a, t)).tobytes())


def read_mesh_binary(path: str) -> dict:
    with open(path, "rb") as f:
        if f.read(8) != MESH_MAGIC:
            raise ValueError(f"{path}: not a cfdb binary mesh")
        cnt = np.frombuffer(f.read(48), I32)
        npoin, nelem, nfixrho, nfixvi, nfixv, nwall, nfixt, nsets, nmaster, nslave, nfix_move, nmove = (int(c) for c in cnt)

        def rd(n, t, shape=None):
            a = np.frombuffer(f.read(n * np.dtype(t).itemsize), t).copy()
            if a.size != n:
                raise ValueError(f"{path}: truncated")
            return a.reshape(shape) if shape else a

        out = dict(X=rd(npoin, F64), Y=rd(npoin, F64), inpoel=rd(3 * nelem, I32, (nelem, 3)))
        out["fixrho"] = (rd(nfixrho, I32), rd(nfixrho, F64))
        out["fixvi"] = (rd(nfixvi, I32), rd(nfixvi, F64), rd(nfixvi, F64))
        out["fixv"] = rd(nfixv, I32)
        out["wall"] = rd(2 * nwall, I32, (nwall, 2))
        out["fixt"] = (rd(nfixt, I32), rd(nfixt, F64))
        out["sets"] = rd(4 * nsets, I32, (nsets, 4))
        out["master"], out["slave"], out["ifm"], out["i_m"] = rd(nmaster, I32), rd(nslave, I32), rd(nfix_move, I32), rd(nmove, I32)
    return out


def write_deck(raw: RawCase, directory: str, binary_mesh: bool = False) -> None:
    """Write EULER.DAT, <name>-1.dat, <name>.dat in the layout dataLoader.f90 reads (binary_mesh: <name>.cfdbmesh instead of
    the text mesh file; read_deck and host/deck_reader.h prefer it when present)."""
    os.makedirs(directory, exist_ok=True)
    if binary_mesh:
        write_mesh_binary(raw, os.path.join(directory, raw.name + ".cfdbmesh"))
    with open(os.path.join(directory, "EULER.DAT"), "w") as f:
        f.write(raw.name + "\n")
    with open(os.path.join(directory, raw.name + "-1.dat"), "w") as f:
        w = f.write
        w("IRESTART MAXITER IPRINT MOVIE ITLOCAL\n")
        w(f"{raw.IRESTART} {raw.MAXITER} {raw.IPRINT} {raw.MOVIE} {raw.ITLOCAL}\n")
        w("FSAFE U_inf V_inf MACH_inf T_inf RHO_inf P_inf\n")
        w(" ".join(_r(v) for v in (raw.FSAFE, raw.U_inf, raw.V_inf, raw.MACH_inf, raw.T_inf, raw.RHO_inf, raw.P_inf)) + "\n")
        w("FMU FGX FGY QH\n")
        w(" ".join(_r(v) for v in (raw.FMU, raw.FGX, raw.FGY, raw.QH)) + "\n")
        w("FK FR FCv GAMA NGAS\n")
        w(" ".join(_r(v) for v in (raw.FK, raw.FR, raw.FCv, raw.GAMA)) + f" {raw.NGAS}\n")
        w("CTE\n")
        w(_r(raw.CTE) + "\n")
        w("--\nMOVING XREF YREF\n")
        w(f"{raw.MOVING} {_r(raw.XREF1)} {_r(raw.YREF1)}\n")
        w("--\nprint flags\n")
        w(" ".join(["'.si.'"] * 7) + "\n")
        w("--\n--\n--\nETA_REFIN HHMAX_REFIN HHMIN_REFIN\n")
        w("0.0 0.0 0.0\n")
    if binary_mesh:
        return
    with open(os.path.join(directory, raw.name + ".dat"), "w") as f:
        w = f.write
        w("NPOIN NELEM\n")
        w(f"{raw.npoin} {raw.nelem}\n")
        w("nfixrho nfixvi nfixv nwall nfixt nsets nmaster nslave nfix_move nmove\n")
        w(" ".join(str(int(v)) for v in (
            len(raw.fixrho[0]), len(raw.fixvi[0]), len(raw.fixv), len(raw.wall), len(raw.fixt[0]), len(raw.sets),
            len(raw.master), len(raw.slave), len(raw.ifm), len(raw.i_m))) + "\n")
        w("--\n--\n--\nCOORDINATES\n")
        for i in range(raw.npoin):
            w(f"{i + 1} {_r(raw.X[i])} {_r(raw.Y[i])}\n")
        w("ELEMENTS\n")
        for i in range(raw.nelem):
            a, b, c = raw.inpoel[i]
            w(f"{i + 1} {a} {b} {c}\n")
        w("FIX RHO\n")
        for n, v in zip(*raw.fixrho):
            w(f"{n} {_r(v)}\n")
        w("FIX VEL INFLOW\n")
        for n, a, b in zip(*raw.fixvi):
            w(f"{n} {_r(a)} {_r(b)}\n")
        w("FIX VEL NO SLIP\n")
        for n in raw.fixv:
            w(f"{n} 0.0 0.0\n")
        w("WALL\n")
        for a, b in raw.wall:
            w(f"{a} {b}\n")
        w("FIX T\n")
        for n, v in zip(*raw.fixt):
            w(f"{n} {_r(v)}\n")
        w("SETS\n")
        for e, a, b, s in raw.sets:
            w(f"{e} {a} {b} {s}\n")
        w("MASTER\n")
        for n in raw.master:
            w(f"{n}\n")
        w("SLAVE\n")
        for n in raw.slave:
            w(f"{n}\n")
        w("FIX MOVE\n")
        for n in raw.ifm:
            w(f"{n} 0.0\n")
        w("MOVE\n")
        for n in raw.i_m:
            w(f"{n} 0.0\n")


class _Lines:
    def __init__(self, path):
        with open(path) as f:
            self.l = f.read().split("\n")
        self.i = 0

    def skip(self, n=1):
        self.i += n

    def vals(self):
        s = self.l[self.i].replace(",", " ").split()
        self.i += 1
        return s

    def block(self, n, ncol):
        rows = [self.vals()[:ncol] for _ in range(n)]
        return rows


def _fnum(s):  # list-directed reals accept Fortran 'd' exponents
    return float(s.replace("d", "e").replace("D", "e"))


def read_deck(directory: str) -> RawCase:
    """Inverse of write_deck; follows the read sequence of dataLoader.f90:22-56 and :98-255."""
    with open(os.path.join(directory, "EULER.DAT")) as f:
        name = f.readline().strip()
    L = _Lines(os.path.join(directory, name + "-1.dat"))
    L.skip(); v = L.vals(); IRESTART, MAXITER, IPRINT, MOVIE, ITLOCAL = (int(x) for x in v[:5])
    L.skip(); v = L.vals(); FSAFE, U_inf, V_inf, MACH_inf, T_inf, RHO_inf, P_inf = (_fnum(x) for x in v[:7])
    L.skip(); v = L.vals(); FMU, FGX, FGY, QH = (_fnum(x) for x in v[:4])
    L.skip(); v = L.vals(); FK, FR, FCv, GAMA = (_fnum(x) for x in v[:4]); NGAS = int(v[4])
    L.skip(); CTE = _fnum(L.vals()[0])
    L.skip(2); v = L.vals(); MOVING = int(v[0]); XREF1, YREF1 = _fnum(v[1]), _fnum(v[2])
    common = dict(name=name, IRESTART=IRESTART, MAXITER=MAXITER, IPRINT=IPRINT, MOVIE=MOVIE,
                  ITLOCAL=ITLOCAL, FSAFE=FSAFE, U_inf=U_inf, V_inf=V_inf, MACH_inf=MACH_inf, T_inf=T_inf, RHO_inf=RHO_inf,
                  P_inf=P_inf, FMU=FMU, FGX=FGX, FGY=FGY, QH=QH, FK=FK, FR=FR, FCv=FCv, GAMA=GAMA, NGAS=NGAS, CTE=CTE,
                  MOVING=MOVING, XREF1=XREF1, YREF1=YREF1)
    bin_path = os.path.join(directory, name + ".cfdbmesh")
    if os.path.exists(bin_path):
        return RawCase(**common, **read_mesh_binary(bin_path))
    M = _Lines(os.path.join(directory, name + ".dat"))
    M.skip(); npoin, nelem = (int(x) for x in M.vals()[:2])
    M.skip(); cnt = [int(x) for x in M.vals()[:10]]
    nfixrho, nfixvi, nfixv, nwall, nfixt, nsets, nmaster, nslave, nfix_move, nmove = cnt
    M.skip(4)
    X = np.zeros(npoin); Y = np.zeros(npoin)
    for r in M.block(npoin, 3):
        i = int(r[0]) - 1; X[i] = _fnum(r[1]); Y[i] = _fnum(r[2])
    M.skip()
    inpoel = np.zeros((nelem, 3), I32)
    for r in M.block(nelem, 4):
        inpoel[int(r[0]) - 1] = [int(r[1]), int(r[2]), int(r[3])]
    M.skip(); b = M.block(nfixrho, 2)
    fixrho = (_i([int(r[0]) for r in b]), _f([_fnum(r[1]) for r in b]))
    M.skip(); b = M.block(nfixvi, 3)
    fixvi = (_i([int(r[0]) for r in b]), _f([_fnum(r[1]) for r in b]), _f([_fnum(r[2]) for r in b]))
    M.skip(); b = M.block(nfixv, 3)
    fixv = _i([int(r[0]) for r in b])
    M.skip(); b = M.block(nwall, 2)
    wall = _i([[int(r[0]), int(r[1])] for r in b], 2)
    M.skip(); b = M.block(nfixt, 2)
    fixt = (_i([int(r[0]) for r in b]), _f([_fnum(r[1]) for r in b]))
    M.skip(); b = M.block(nsets, 4)
    sets = _i([[int(x) for x in r] for r in b], 4)
    M.skip(); master = _i([int(r[0]) for r in M.block(nmaster, 1)])
    M.skip(); slave = _i([int(r[0]) for r in M.block(nslave, 1)])
    M.skip(); ifm = _i([int(r[0]) for r in M.block(nfix_move, 1)])
    M.skip(); i_m = _i([int(r[0]) for r in M.block(nmove, 1)])
    return RawCase(
        name=name, X=X, Y=Y, inpoel=inpoel, IRESTART=IRESTART, MAXITER=MAXITER, IPRINT=IPRINT, MOVIE=MOVIE,
        ITLOCAL=ITLOCAL, FSAFE=FSAFE, U_inf=U_inf, V_inf=V_inf, MACH_inf=MACH_inf, T_inf=T_inf, RHO_inf=RHO_inf,
        P_inf=P_inf, FMU=FMU, FGX=FGX, FGY=FGY, QH=QH, FK=FK, FR=FR, FCv=FCv, GAMA=GAMA, NGAS=NGAS, CTE=CTE,
        MOVING=MOVING, XREF1=XREF1, YREF1=YREF1, fixrho=fixrho, fixvi=fixvi, fixv=fixv, wall=wall, fixt=fixt,
        sets=sets, master=master, slave=slave, ifm=ifm, i_m=i_m,
    )


# ---------------------------------------------------------------------------------------------
# restart file <name>.RST — Fortran unformatted sequential (PRINTREST ns2DComp.ALE.f90:898-917, RESTART :423-431):
# record 1 = (ITER int32, TIME real64); then one record per node = (U(1:4), T, GAMM) real64.  Each record is
# framed by 4-byte length markers (gfortran / ifort default).

def write_rst(path: str, it: int, time: float, U, T, GAMM) -> None:
    import struct

    U = np.asarray(U, F64).reshape(-1, 4)
    n = U.shape[0]
    rec = np.zeros(n, dtype=[("l0", "<i4"), ("u", "<f8", 4), ("t", "<f8"), ("g", "<f8"), ("l1", "<i4")])
    rec["l0"] = rec["l1"] = 48
    rec["u"], rec["t"], rec["g"] = U, np.asarray(T, F64), np.asarray(GAMM, F64)
    with open(path, "wb") as f:
        f.write(struct.pack("<iidi", 12, int(it), float(time), 12))
        f.write(rec.tobytes())


def read_rst(path: str, npoin: int):
    """Returns (ITER, TIME, U(npoin,4), T, GAMM).  The reference reads ITER/TIME into locals and drops them, and does
    not restore VEL_X/VEL_Y; callers decide what to do with them."""
    import struct

    with open(path, "rb") as f:
        l0, it, time, l1 = struct.unpack("<iidi", f.read(20))
        if l0 != 12 or l1 != 12:
            raise ValueError("not a Fortran unformatted .RST file (bad first record)")
        rec = np.frombuffer(f.read(npoin * 56), dtype=[("l0", "<i4"), ("u", "<f8", 4), ("t", "<f8"), ("g", "<f8"), ("l1", "<i4")])
    if rec.size != npoin or (rec["l0"] != 48).any() or (rec["l1"] != 48).any():
        raise ValueError("restart file does not hold npoin node records")
    return it, time, np.ascontiguousarray(rec["u"]), np.ascontiguousarray(rec["t"]), np.ascontiguousarray(rec["g"])
