#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== sanity"; timeout 150 python tests/opt_worker.py 2>&1 | tail -1
for lib in "" _ab_re152; do
  echo "== lib $lib"
  CFDB_LIB_PATH=$PWD/cfd_b200/libcfdb200$lib.so CFDB_STAGE_STATS=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep -E "stage_fused|Error|error" | tail -2 | cut -c1-330 | tee -a gpurun_out/r2_exp9.txt
done
CFDB_STAGE_STATS=1 timeout 300 python tools/exp_stage.py 2829 visc 2>&1 | grep -E "stage_fused|Error|error" | tail -2 | cut -c1-330 | tee -a gpurun_out/r2_exp9.txt
