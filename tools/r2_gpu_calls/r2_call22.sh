#!/bin/bash
# stage_fused cycle counters with the whole barrier wait timed
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for w in 1 2; do
  CFDB_STAGE_STATS=$w timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "cycles per tile\|default\|STATS" | tail -2
done
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k spmv 2>&1 | tail -2
