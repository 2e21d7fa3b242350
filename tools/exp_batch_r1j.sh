#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/exp_r1j.txt
: > $OUT
run() { echo "== $1" >> $OUT; shift; env "$@" timeout 300 python tools/exp_stage.py 2829 >> $OUT 2>&1; }
run default X=1
run nopre CFDB_BICG_NOPRE=1
run default X=1
run nopre CFDB_BICG_NOPRE=1
echo "== ale default" >> $OUT; timeout 300 python tools/exp_ale.py >> $OUT 2>&1
echo "== ale nopre" >> $OUT; CFDB_BICG_NOPRE=1 timeout 300 python tools/exp_ale.py >> $OUT 2>&1
cut -c1-260 $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1j_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/r1j_pytest.log | tail -3
