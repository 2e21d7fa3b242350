"""Runs three small cases on the GPU and compares with the oracle bit for bit; launched by
tests/test_gpu_parity.py::test_optional_paths_bit_exact in a subprocess so that the opt-in code paths, which the
library selects from environment variables read once per process, can each be exercised."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cfd_b200 import deck, meshgen  # noqa: E402
from cfd_b200.solver import NSComp2D  # noqa: E402
from oracle.orclib import Oracle  # noqa: E402

CASES = {
    "euler": (lambda: meshgen.channel(nx=41, ny=13), True),
    "visc_itlocal": (lambda: meshgen.channel(nx=37, ny=11, FMU=1.8e-5, FK=0.0257, ITLOCAL=40), True),
    "ale": (lambda: meshgen.ale_body(nt=48, nr=14), False),
}
FIELDS = ["U", "U1", "RHS", "T", "VEL_X", "VEL_Y", "P", "RMACH", "SHOC", "T_SUGN2", "X", "Y", "W_X", "M"]
for name, (mk, bump) in CASES.items():
    lc = deck.load(mk())
    g, o = NSComp2D(lc), Oracle(lc)
    if bump:
        for k, v in meshgen.density_bump(lc).items():
            g.set(k, v)
            o.set(k, v)
    g.step(6)
    o.step(6)
    for f in FIELDS:
        a, b = g.get(f), o.get(f)
        if not np.array_equal(a.view(np.uint64), b.view(np.uint64)):
            raise SystemExit(f"{name}: {f} differs ({int((a.view(np.uint64) != b.view(np.uint64)).sum())} entries)")
    assert g.scalar("bicg_y") == o.scalar("bicg_y") and g.scalar("DTMIN") == o.scalar("DTMIN")
import pathological  # noqa: E402  (tests/ is sys.path[0] for this script)

lc = deck.load(meshgen.channel(nx=41, ny=13, FMU=1.8e-5, FK=0.0257))
pathological.check(lc, NSComp2D(lc), Oracle(lc))
print("OPT_PATH_OK")
