#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
CFDB_VERBOSE=1 timeout 300 python tools/exp_stage.py 1415 visc 2>&1 | grep "fallbacks\|ms/step" | tail -3
CFDB_VERBOSE=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "fallbacks\|ms/step" | tail -3
