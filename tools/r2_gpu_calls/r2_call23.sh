#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
CFDB_STAGE_STATS=2 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "stage_fused\]" | tail -2
