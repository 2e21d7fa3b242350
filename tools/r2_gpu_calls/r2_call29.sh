#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
CFDB_VERBOSE=1 timeout 600 python tools/exp_ale.py 2>&1 | grep "fallbacks\|ms/step" | tail -4
CFDB_VERBOSE=1 timeout 300 python tools/exp_stage.py 1415 2>&1 | grep "fallbacks\|ms/step" | tail -3
