#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for env in "X=1" "CFDB_BNODE_TILE_ORDER=1"; do
  env $env timeout 300 python tools/exp_stage.py 2829 2>&1 | grep -E "stage_fused|Error|error" | tail -1 | cut -c1-300 | tee -a gpurun_out/r2_exp11.txt
done
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gputests11.log; cat gpurun_out/r2_gputests11.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; tail -c 6000 gpurun_out/r2_bench_n1_b.json; tail -5 gpurun_out/r2_bench_n1_b.err
