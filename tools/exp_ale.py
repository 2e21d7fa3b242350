"""GPU experiment helper: per-kernel event timings of the moving-mesh step (BASELINE configs[2]: O-mesh around a pitching ellipse)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_b200 import deck, meshgen  # noqa: E402
from cfd_b200.solver import NSComp2D  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
lc = deck.load(meshgen.ale_body(nt=nt, nr=nr, IPRINT=10**9, MAXITER=10**9))
g = NSComp2D(lc, use_gcl=1)
g.step(3)
g.sync()
t0 = time.perf_counter()
g.step(10)
g.sync()
dt = (time.perf_counter() - t0) / 10
g.profile(True)
g.step(3)
g.sync()
out = [f"E={lc.nelem} ms/step={dt*1e3:.3f} (wall, 10 steps in one call) bicg iters {g.scalar('bicg_x'):.0f}/{g.scalar('bicg_y'):.0f}"]
tot = 0.0
for kn in ("deriv", "masas", "normales", "deltat", "dt_logic", "dtl", "estab", "calcrhs_elem", "node_update", "dot", "norms", "spmv", "vec",
           "fixrows", "scalar", "laplace", "transf", "move_apply", "forces", "gcl", "layout", "fill", "halo", "stage_fused"):
    try:
        ms, cnt = g.profile_get(kn)
    except Exception:
        continue
    if cnt:
        out.append(f"{kn}={ms/cnt:.4f}ms x{cnt/3:.1f}")
        tot += ms / 3
out.append(f"sum={tot:.3f}ms/step")
print(" ".join(out), flush=True)
