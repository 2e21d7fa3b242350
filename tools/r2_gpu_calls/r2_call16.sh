#!/bin/bash
# A/B: register split of the stage kernel, node chain form, tile size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for t in _re152 _re136 _nodeplain; do
  echo "== $t"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab$t.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
done
echo "== TE512"; CFDB_TILE_TE=512 CFDB_VERBOSE=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -3
echo "== default"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
