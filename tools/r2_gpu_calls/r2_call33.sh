#!/bin/bash
# bisect of the stage kernel's 5 % slow-down across the counter commits
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for t in 0e2be24 7d786e3 da2de3c; do
  echo "== $t"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_$t.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-160
done
echo "== current"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-160
echo "== 0e2be24 again"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_0e2be24.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-160
