#!/bin/bash
# node warps: contributions per batch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for t in _nb6 _nb7; do echo "== $t"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab$t.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-140; done
echo "== 8"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-140
echo "== _nb6 again"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_nb6.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-140
