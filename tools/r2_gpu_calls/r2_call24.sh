#!/bin/bash
# early hand-back of the stream slot + C-free wait moved to the stores: parity, timing, counters
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c24_tests.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_c24_tests.txt
timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
CFDB_STAGE_STATS=2 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "stage_fused\]" | tail -2
