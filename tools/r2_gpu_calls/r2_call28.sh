#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
ALE=1 timeout 300 python tools/exp_stage.py 1415 2>&1 | tail -1
timeout 300 python tools/exp_stage.py 1415 2>&1 | tail -1
timeout 600 python tools/exp_ale.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ale or forces" 2>&1 | tail -2
