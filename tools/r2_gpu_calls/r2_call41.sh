#!/bin/bash
# parallel tile construction + double-buffered streamed staging: parity suite, create phases, the bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c41_tests.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_c41_tests.txt
CFDB_VERBOSE=1 timeout 300 python tools/exp_stage.py 2829 2>&1 | grep "cfdb_create\|ms/step" | cut -c1-200 | tail -12
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
