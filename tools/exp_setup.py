"""GPU experiment helper: time cfdb_create (+init) on the bench mesh, device-built vs host-built topology."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_b200 import deck, meshgen  # noqa: E402
from cfd_b200.solver import NSComp2D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2829
t0 = time.perf_counter()
lc = deck.load(meshgen.square(n=n, IPRINT=10**9, MAXITER=10**9))
t1 = time.perf_counter()
g = NSComp2D(lc, init=False)
g.sync()
t2 = time.perf_counter()
g.L.cfdb_init(g.h)
g.sync()
t3 = time.perf_counter()
print(f"host_topo={os.environ.get('CFDB_HOST_TOPO', '0')} E={lc.nelem} meshgen={t1 - t0:.2f}s cfdb_create={t2 - t1:.2f}s cfdb_init={t3 - t2:.2f}s", flush=True)
