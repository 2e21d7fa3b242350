#!/bin/bash
# round 2, GPU call 4: decoupled fused stage -- sanity, stage statistics, timings at three tile sizes, one ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for te in 352 384 416; do
  echo "== sanity TE=$te"; CFDB_TILE_TE=$te timeout 150 python tests/opt_worker.py 2>&1 | tail -3
done
echo "== timings"
for env in "CFDB_TILE_TE=352" "CFDB_TILE_TE=384" "CFDB_TILE_TE=416"; do
  env $env CFDB_STAGE_STATS=1 CFDB_VERBOSE=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep -E "stage_fused|tiles|ms/step|stage tiles" | tail -8 | tee -a gpurun_out/r2_exp4.txt
done
env CFDB_TILE_TE=352 timeout 300 python tools/exp_stage.py 2829 visc 2>&1 | tail -1 | tee -a gpurun_out/r2_exp4.txt
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_fused -s 6 -c 1 -o gpurun_out/r2_fused -f python tools/exp_stage.py 2829 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
