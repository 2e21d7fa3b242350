#!/bin/bash
# round 2, GPU call 3: first run of the fused tile stage
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== sanity (3 small cases vs oracle)"; timeout 180 python tests/opt_worker.py 2>&1 | tail -5
rc=$?; echo "sanity rc=$rc"
echo "== sanity TE=384"; CFDB_TILE_TE=384 timeout 180 python tests/opt_worker.py 2>&1 | tail -3
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gputests3.log; cat gpurun_out/r2_gputests3.log
echo "== timings"
for env in "" "CFDB_TILE_TE=384" "CFDB_NO_FUSED=1" "CFDB_NO_PERM=1 CFDB_NO_FUSED=1"; do
  env $env timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | tee -a gpurun_out/r2_exp3.txt
done
for env in "" "CFDB_TILE_TE=384" "CFDB_NO_FUSED=1"; do
  env $env timeout 300 python tools/exp_stage.py 2829 visc 2>&1 | tail -1 | tee -a gpurun_out/r2_exp3.txt
done
echo "== memcheck (small)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/opt_worker.py 2>&1 | tail -6
