#!/bin/bash
# nodal records + RCB tiles: parity suite, then per-kernel timings (RCB vs Morton tiles), then a short bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c14_tests.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_c14_tests.txt
timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
CFDB_TILE_ORDER=morton timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
CFDB_VERBOSE=1 timeout 300 python tools/exp_stage.py 2829 visc 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_c14_bench.json 2> gpurun_out/r2_c14_bench.err; cut -c1-1500 gpurun_out/r2_c14_bench.json
