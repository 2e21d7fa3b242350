#!/bin/bash
# boundary records + cheaper waits: parity suite, per-kernel timings, stage statistics
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c15_tests.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_c15_tests.txt
timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
CFDB_STAGE_STATS=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -4
CFDB_FUSED_VISC=1 timeout 300 python tools/exp_stage.py 2829 visc 2>&1 | tail -1
