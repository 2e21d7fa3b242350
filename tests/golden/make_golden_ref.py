"""Generate tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN FORTRAN SOURCES (read from /root/reference at
generation time, never copied) with oracle/f90ref -- a Fortran-subset translator written for this purpose, because no
Fortran compiler exists in the image.  What runs is `PROGRAM NSComp2D` itself: readInputData and
loadMeshData on a deck written by cfd_b200.deck.write_deck, RESTART (free stream), smoothing, NORMALES, DERIV, MASAS,
laplace, then MAXITER passes of the time loop (DELTAT, the dt logic, RK with CUARTO_ORDEN/ESTAB/calcRHS/FUENTE/
FIXVEL/NORMALVEL/FIX, fluidStructure with FORCES/TRANSF/biCG, the residual norms, the MOVING geometry refresh).

These vectors are what pins the oracle (and through it the CUDA path) to the reference:
    python tests/golden/make_golden_ref.py          # needs /root/reference; a few seconds per case

Two things are injected, both stated in DESIGN.md section 3:
  * an initial density bump, written into U/T right after RESTART returns (the reference starts from uniform free stream;
    the bump makes every term of calcRHS non-trivial from step 1);
  * for the moving-mesh case only, biCG's `vecdot` is evaluated in the build's canonical summation order: it is the one
    place where the reference leaves the order to OpenMP (`reduction(+:res)`, biconjGrad.f90:162).  `ref_ale_seqdot.npz`
    keeps the same run with the reference's own sequential loop, to show the difference is round-off.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from cfd_b200 import deck, meshgen  # noqa: E402


def _noslip():
    """viscous channel with a no-slip lower wall (single-precision TWALL, F11), a fixed-T patch, duplicated list entries"""
    nx, ny = 21, 9
    raw = meshgen.channel(nx=nx, ny=ny, FMU=1.8e-5, FK=0.0257, mach=0.6)
    lower = np.arange(2, nx, dtype=np.int32)
    raw.wall = raw.wall[~np.isin(raw.wall, lower[4:]).any(1)]
    raw.fixv = np.concatenate([lower, lower[:3]]).astype(np.int32)
    top = ((ny - 1) * nx + np.arange(4, 10)).astype(np.int32)
    raw.fixt = (np.concatenate([top, top[:2]]).astype(np.int32), np.concatenate([np.full(6, 1.1), np.full(2, 0.9)]))
    raw.fixrho = (np.concatenate([raw.fixrho[0], raw.fixrho[0][:2]]).astype(np.int32),
                  np.concatenate([raw.fixrho[1], np.array([-1.0, 1.05])]))
    return raw


# name -> (RawCase factory, steps, density bump?, canonical vecdot?)
CASES = {
    "ref_channel_visc": (lambda: meshgen.channel(nx=21, ny=9, FMU=1.8e-5, FK=0.0257), 6, True, False),
    "ref_channel_euler_itlocal": (lambda: meshgen.channel(nx=17, ny=9, ITLOCAL=5), 8, True, False),
    "ref_channel_noslip": (_noslip, 6, True, False),
    "ref_wedge": (lambda: meshgen.wedge(nx=25, ny=13, mach=2.5), 6, True, False),
    "ref_ale": (lambda: meshgen.ale_body(nt=32, nr=8), 5, False, True),
    "ref_ale_seqdot": (lambda: meshgen.ale_body(nt=32, nr=8), 5, False, False),
    # viscous moving mesh: FUENTE with W /= 0 next to the viscous terms, FORCES and FORCE_VISC on the body set
    "ref_ale_visc": (lambda: meshgen.ale_body(nt=32, nr=8, FMU=1.8e-5, FK=0.0257), 4, False, True),
    # 1000 passes of the reference's time loop (about five minutes of interpretation): north_star's long-run tolerance
    "ref_channel_1000": (lambda: meshgen.channel(nx=17, ny=9, FMU=1.8e-5, FK=0.0257), 1000, True, False),
    # BASELINE configs[0]'s mesh itself: 10 251 nodes / 20 000 triangles, smoothing included, three passes of the loop
    # (a minute of interpretation per pass); only the arrays in SUBSET are kept, to bound the file size
    "ref_channel_10k": (lambda: meshgen.channel(nx=201, ny=51, FMU=1.8e-5, FK=0.0257), 3, True, False),
}
IPRINT = {"ref_channel_1000": 100}
SUBSET = {"ref_channel_10k": ["U", "RHS", "T", "X", "Y", "M", "SHOC", "T_SUGN2", "esup2", "psup2", "lap_rowptr", "n_m", "n_ipoin",
                              "n_x", "n_y", "cnv", "dtmin", "time"]}
NODE_FIELDS = ["T", "P", "RHO", "E", "RMACH", "VEL_X", "VEL_Y", "W_X", "W_Y", "X", "Y", "M"]
ELEM_FIELDS = ["SHOC", "T_SUGN1", "T_SUGN2", "T_SUGN3", "area"]
FIELDS = ["U", "RHS"] + NODE_FIELDS + ELEM_FIELDS + ["dNx", "dNy", "lap_sparse"]
INT_FIELDS = ["esup1", "esup2", "psup1", "psup2", "lap_idx", "lap_rowptr"]


def raw_case(name):
    raw = CASES[name][0]()
    raw.IPRINT = IPRINT.get(name, 1)          # residual norms every IPRINT steps
    raw.MAXITER = CASES[name][1]
    return raw


def run_reference(name):
    """-> dict of arrays in the repo's layouts (Fortran memory order: U as [npoin][4], dNx as [nelem][3])"""
    from oracle import orclib
    from oracle.f90ref.refrun import Reference

    raw = raw_case(name)
    _, steps, bump, canon = CASES[name]
    lc0 = deck.load(raw)
    st = meshgen.density_bump(lc0) if bump else None
    trace = {"dtmin": [], "time": [], "bicg_calls": 0}

    def hook(r):
        ns = r.ns
        fs = ns["p_meshmove__fluidstructure"]

        def fs_traced(dtmin, time, *a):
            trace["dtmin"].append(float(dtmin))
            trace["time"].append(float(time))
            return fs(dtmin, time, *a)

        ns["p_meshmove__fluidstructure"] = fs_traced

    ref = Reference()
    cnv = ref.run_program(raw, hook=hook, initial_state=st, canonical_vecdot=canon)
    g, v, md = ref.mod("mvariabgen"), ref.mod("mvariables"), ref.mod("meshdata")
    vel, est, lap, pn = ref.mod("mvelocidades"), ref.mod("mestabilizacion"), ref.mod("mlaplace"), ref.mod("pointneighbor")
    nor, mm = ref.mod("mnormales"), ref.mod("meshmove")
    nedges = int(md.nsets) if float(ref.mod("inputdata").fmu) != 0.0 else 0
    out = {
        "U": g.u.T.ravel(), "RHS": g.rhs.T.ravel(),
        "T": v.t, "P": v.p, "RHO": v.rho, "E": v.e, "RMACH": v.rmach,
        "VEL_X": vel.vel_x, "VEL_Y": vel.vel_y, "W_X": vel.w_x, "W_Y": vel.w_y,
        "X": md.x, "Y": md.y, "M": md.m, "area": md.area, "dNx": md.dnx.ravel(order="F"), "dNy": md.dny.ravel(order="F"),
        "SHOC": est.shoc, "T_SUGN1": est.t_sugn1, "T_SUGN2": est.t_sugn2, "T_SUGN3": est.t_sugn3,
        "lap_sparse": lap.lap_sparse, "lap_idx": lap.lap_idx, "lap_rowptr": lap.lap_rowptr,
        "esup1": pn.esup1, "esup2": pn.esup2, "psup1": pn.psup1, "psup2": pn.psup2,
        "n_m": np.array([nor.m]), "n_ipoin": nor.n_ipoin[:nor.m], "n_x": nor.n_x[:nor.m], "n_y": nor.n_y[:nor.m],
        "cnv": np.array([[float(x) for x in rec] for rec in cnv]),
        "FX": mm.fx, "FY": mm.fy, "RM": mm.rm, "F_VX": mm.f_vx, "F_VY": mm.f_vy,
        "skin": np.array([[float(x) for x in rec] for rec in ref.io.written.get("SKIN.DAT", [])[-nedges:]]).reshape(-1, 3),
        "dtmin": np.array(trace["dtmin"]), "time": np.array(trace["time"]),
        # what readInputData / loadMeshData left in the modules (the deck-reader boundary)
        "in_ifixv_node": md.ifixv_node, "in_rfixv_valuex": md.rfixv_valuex, "in_rfixv_valuey": md.rfixv_valuey,
        "in_ifixrho_node": md.ifixrho_node, "in_rfixrho_value": md.rfixrho_value,
        "in_ifixt_node": md.ifixt_node, "in_rfixt_value": md.rfixt_value, "in_ilaux": md.ilaux if md.ilaux is not None else np.zeros(0, np.int32),
    }
    if name == "ref_ale_visc":   # PRINTFLAVIA's GiD file of the last print step (MOVIE = 0: rewritten every time), byte for byte
        with open(os.path.join(HERE, name + ".flavia.res"), "w") as f:
            f.write("\n".join(ref.io.text[raw.name + ".flavia.res"]) + "\n")
    if name in SUBSET:
        out = {k: out[k] for k in SUBSET[name]}
    return {k: np.array(a) for k, a in out.items()}


def main():
    from oracle.f90ref import runtime as rt
    for name in (sys.argv[1:] or CASES):
        out = run_reference(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "npoin", out["X"].size, "nelem", out["SHOC"].size, "steps", len(out["dtmin"]), "cnv[-1]", out["cnv"][-1])
    print("x**1.5/.5/-.5 evaluations:", rt.pow_stats, "(glibc pow differs from the correctly rounded value in that many)")


if __name__ == "__main__":
    main()
