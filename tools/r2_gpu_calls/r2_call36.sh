#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
echo "== noecb (timing only)"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_noecb.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-200
CFDB_STAGE_STATS=2 CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_noecb.so timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "stage_fused\]" | tail -2
echo "== current"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-200
