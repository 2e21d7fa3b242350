#!/bin/bash
# forces by one CTA per body set: parity suite, moving-mesh step breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c27_tests.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_c27_tests.txt
timeout 600 python tools/exp_ale.py 2>&1 | tail -1
