#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cat > /tmp/san_worker.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
from cfd_b200 import deck, meshgen, capi
from cfd_b200.solver import NSComp2D
lc = deck.load(meshgen.ale_body(nt=32, nr=8, FMU=1.8e-5, FK=0.0257, IPRINT=1, MAXITER=3))
capi.smoothing(lc)
g = NSComp2D(lc)
g.step(3)
print(g.force_visc()[0][:2], g.get("esup1")[:5], g.get("psup1")[:5], g.get("lap_idx")[:5])
g.printflavia("/tmp/x.flavia.res", 3)
print("SAN_OK")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san_worker.py > gpurun_out/memcheck_r1c.log 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck_r1c.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tests/opt_worker.py > gpurun_out/memcheck_r1c_opt.log 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck_r1c_opt.log
tail -4 gpurun_out/memcheck_r1c.log; tail -4 gpurun_out/memcheck_r1c_opt.log
