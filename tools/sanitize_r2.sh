#!/bin/bash
# compute-sanitizer over the round-2 kernels on small cases: the fused stage + boundary_update (fixed mesh, graph replay and
# stream launches, 3 ranks of tiles so that the rings wrap), the moving-mesh step (forces, estab fast path), the streamed step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/san2_worker.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from cfd_b200 import deck, meshgen
from cfd_b200.solver import NSComp2D
lc = deck.load(meshgen.square(n=61, IPRINT=10**9, MAXITER=10**9))        # 7 200 triangles = 19 tiles
g = NSComp2D(lc)
for k, v in meshgen.density_bump(lc).items():
    g.set(k, v)
g.step(4)
g.sync()
print("fixed", float(np.abs(g.get("U")).sum()), g.scalar("tile_interior"))
la = deck.load(meshgen.ale_body(nt=48, nr=12, IPRINT=10**9, MAXITER=10**9))
a = NSComp2D(la, use_gcl=1)
a.step(3)
a.sync()
print("ale", float(np.abs(a.get("U")).sum()), a.scalar("FX1"))
print("SAN_OK")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san2_worker.py > gpurun_out/r2_$tool.log 2>&1; echo "$tool exit $?" >> gpurun_out/r2_$tool.log
  tail -5 gpurun_out/r2_$tool.log
done
CFDB_NO_GRAPH=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san2_worker.py > gpurun_out/r2_memcheck_nograph.log 2>&1; echo "memcheck(no graph) exit $?" >> gpurun_out/r2_memcheck_nograph.log; tail -3 gpurun_out/r2_memcheck_nograph.log
