#!/bin/bash
# the element stream completing on the tile's inputs-landed barrier (one wait fewer per tile and element warp)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-140
timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-140
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c43_tests.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_c43_tests.txt
