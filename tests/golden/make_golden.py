"""Generate tests/golden/*.npz from the ORACLE (oracle/oracle.cpp), this container, seeded inputs.

The reference ships no golden vectors and cannot be built here, so these vectors do not
pin the oracle to the Fortran program; they freeze the oracle's own results so that later changes to
oracle.cpp (or to the mesh generator) are caught, and they give the GPU tests a comparison that needs no
oracle at run time.      python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cfd_b200 import deck, meshgen  # noqa: E402
from oracle.orclib import Oracle  # noqa: E402

CASES = {
    "channel_visc": (lambda: meshgen.channel(nx=41, ny=13, FMU=1.8e-5, FK=0.0257), True, 0),
    "wedge": (lambda: meshgen.wedge(nx=49, ny=25, mach=2.5), True, 0),
    "ale_gcl": (lambda: meshgen.ale_body(nt=48, nr=14), False, 1),
}
STEPS = 5
FIELDS = ["U", "T", "RHS", "SHOC", "T_SUGN1", "T_SUGN2", "T_SUGN3", "VEL_X", "VEL_Y", "P", "RMACH", "X", "Y", "M", "W_X",
          "lap_sparse", "area", "dNx"]


def build(name):
    mk, bump, gcl = CASES[name]
    lc = deck.load(mk())
    return lc, bump, gcl


def main():
    for name in CASES:
        lc, bump, gcl = build(name)
        o = Oracle(lc, use_gcl=gcl)
        if bump:
            for k, v in meshgen.density_bump(lc).items():
                o.set(k, v)
        o.step(STEPS)
        out = {f: o.get(f) for f in FIELDS}
        out.update(mesh_X=lc.X, mesh_Y=lc.Y, mesh_inpoel=lc.inpoel, esup1=o.get("esup1"), psup1=o.get("psup1"),
                   lap_idx=o.get("lap_idx"), scalars=np.array([o.scalar(s) for s in ("DTMIN", "TIME", "ITER", "BANDERA", "bicg_x", "bicg_y")]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, lc.npoin, lc.nelem, out["scalars"])


if __name__ == "__main__":
    main()
