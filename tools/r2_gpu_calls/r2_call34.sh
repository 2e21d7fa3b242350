#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
echo "== current"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-160
echo "== nofb"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_nofb.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-160
echo "== 0e2be24"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_0e2be24.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-160
echo "== current"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-160
CFDB_STAGE_STATS=2 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "stage_fused\]" | tail -2
