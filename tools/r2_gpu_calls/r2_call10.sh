#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== sanity"; timeout 150 python tests/opt_worker.py 2>&1 | tail -1
echo "== sanity boundary pass"; CFDB_BOUNDARY_PASS=1 timeout 150 python tests/opt_worker.py 2>&1 | tail -1
for env in "X=1" "CFDB_BOUNDARY_PASS=1"; do
  env $env CFDB_STAGE_STATS=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep -E "stage_fused|Error|error" | tail -2 | cut -c1-360 | tee -a gpurun_out/r2_exp10.txt
done
echo "== pytest"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gputests10.log; cat gpurun_out/r2_gputests10.log
