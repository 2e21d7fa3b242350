#!/bin/bash
# round 2, GPU call 6: warp-specialised fused stage (element warps / node warps / loader)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== sanity"; timeout 150 python tests/opt_worker.py 2>&1 | tail -2
echo "== timings"
CFDB_STAGE_STATS=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep -E "stage_fused|Error|error" | tail -3 | tee -a gpurun_out/r2_exp6.txt
timeout 300 python tools/exp_stage.py 2829 visc 2>&1 | tail -1 | tee -a gpurun_out/r2_exp6.txt
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_fused -s 6 -c 1 -o gpurun_out/r2_fused_ws -f python tools/exp_stage.py 2829 2>&1 | tail -2
