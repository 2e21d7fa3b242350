"""GPU experiment helper: per-kernel event timings of the resident step on the bench mesh."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_b200 import deck, meshgen  # noqa: E402
from cfd_b200.solver import NSComp2D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2829
kw = {}
if len(sys.argv) > 2 and sys.argv[2] == "visc":
    kw = dict(FMU=1.8e-5, FK=0.0257)
lc = deck.load(meshgen.square(n=n, IPRINT=10**9, MAXITER=10**9, **kw))
g = NSComp2D(lc)
if os.environ.get("ALE"):          # moving-mesh kernels (FUENTE, mesh-velocity terms) on the same mesh; W stays 0
    g.set_option("ale", 1)
if os.environ.get("FAST"):
    g.set_option("fast", int(os.environ["FAST"]))
for k, v in meshgen.density_bump(lc).items():
    g.set(k, v)
g.step(3)
g.sync()
t0 = time.perf_counter()
g.step(10)
g.sync()
dt = (time.perf_counter() - t0) / 10
g.profile(True)
g.step(3)
g.sync()
out = [f"E={lc.nelem} ms/step={dt*1e3:.3f}"]
for kn in ("stage_fused", "calcrhs_elem", "node_update", "estab", "deltat", "spmv", "dot", "vec", "fixrows", "scalar", "fill", "dt_logic"):
    ms, cnt = g.profile_get(kn)
    if cnt:
        out.append(f"{kn}={ms/cnt:.4f}ms x{cnt/3:.0f}")
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("CFDB_"))
print(tag or "default", "|", " ".join(out), f"interior={g.scalar('tile_interior'):.3f}", flush=True)
