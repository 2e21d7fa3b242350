#!/bin/bash
# which element warps set the pace of stage_fused (cycle counters of element warps 0..3 = sub-partitions 0..3); polling vs parked waits
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for w in 1 2 3 4; do
  echo "== element warp $((w-1))"; CFDB_STAGE_STATS=$w timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "cycles per tile" | tail -1
done
echo "== espin"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_espin.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
echo "== espin stats w1"; CFDB_STAGE_STATS=2 CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_espin.so timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "cycles per tile" | tail -1
echo "== default"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
