#!/bin/bash
# round 2, GPU call 5: fused stage with branch-free element arithmetic
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for te in 352 384 416; do
  echo "== sanity TE=$te"; CFDB_TILE_TE=$te timeout 150 python tests/opt_worker.py 2>&1 | tail -2
done
echo "== timings"
for env in "CFDB_TILE_TE=352" "CFDB_TILE_TE=384" "CFDB_TILE_TE=416"; do
  env $env CFDB_STAGE_STATS=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep -E "stage_fused|Error|error" | tail -4 | tee -a gpurun_out/r2_exp5.txt
done
env CFDB_TILE_TE=352 timeout 300 python tools/exp_stage.py 2829 visc 2>&1 | tail -1 | tee -a gpurun_out/r2_exp5.txt
env CFDB_TILE_TE=384 timeout 300 python tools/exp_stage.py 2829 visc 2>&1 | tail -1 | tee -a gpurun_out/r2_exp5.txt
echo "== pytest"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gputests5.log; cat gpurun_out/r2_gputests5.log
