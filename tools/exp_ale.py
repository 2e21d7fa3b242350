"""GPU experiment helper: config 3 — pitching-body ALE case (meshMove Laplace biCG + geometry every step)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_b200 import deck, meshgen  # noqa: E402
from cfd_b200.solver import NSComp2D  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
gcl = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lc = deck.load(meshgen.ale_body(nt=nt, nr=nr, IPRINT=10**9, MAXITER=10**9))
g = NSComp2D(lc, use_gcl=gcl)
g.step(2)
g.sync()
t0 = time.perf_counter()
nst = 5
its = []
for _ in range(nst):
    g.step(1)
    its.append((int(g.scalar("bicg_x")), int(g.scalar("bicg_y"))))
g.sync()
dt = (time.perf_counter() - t0) / nst
g.profile(True)
g.step(2)
g.sync()
out = [f"ALE E={lc.nelem} P={lc.npoin} ms/step={dt*1e3:.3f} elem/s={lc.nelem/dt:.3e} bicg_iters={its}"]
for kn in ("calcrhs_elem", "node_update", "estab", "deltat", "spmv", "dot", "vec", "fixrows", "scalar", "laplace", "deriv", "masas", "gcl", "move_apply"):
    ms, cnt = g.profile_get(kn)
    if cnt:
        out.append(f"{kn}={ms/cnt:.4f}ms x{cnt/2:.0f}")
print(" ".join(out), flush=True)
