#!/bin/bash
cd "$GRAFT_REPO_ROOT"
OUT=gpurun_out/exp_r1f.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/exp_stage.py 2829 >> $OUT 2>&1; }
run A=baseline
run CFDB_PF_DIST=75776
run CFDB_PF_DIST=151552
run CFDB_PF_DIST=37888
run CFDB_PF_DIST=303104
run CFDB_PF_DIST=8192
echo "== visc" >> $OUT
python tools/exp_stage.py 2829 visc >> $OUT 2>&1
CFDB_PF_DIST=75776 python tools/exp_stage.py 2829 visc >> $OUT 2>&1
CFDB_PF_DIST=75776 python tests/opt_worker.py >> $OUT 2>&1
cat $OUT | cut -c1-200
