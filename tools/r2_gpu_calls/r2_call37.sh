#!/bin/bash
# loader warp as the fourth node warp: timing against the dedicated loader, counters, parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== merged"; timeout 120 python tools/exp_stage.py 1415 2>&1 | tail -1 | cut -c1-200
echo "== merged"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-200
echo "== separate"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_sep.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-200
CFDB_STAGE_STATS=2 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "stage_fused\]" | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c37_tests.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_c37_tests.txt
