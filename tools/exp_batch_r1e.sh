#!/bin/bash
cd "$GRAFT_REPO_ROOT"
OUT=gpurun_out/exp_r1e.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/exp_stage.py 2829 >> $OUT 2>&1; }
run A=baseline
run CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_hints.so
run A=baseline
run CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_hints.so
echo "== visc" >> $OUT
python tools/exp_stage.py 2829 visc >> $OUT 2>&1
CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_hints.so python tools/exp_stage.py 2829 visc >> $OUT 2>&1
CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_hints.so python tests/opt_worker.py >> $OUT 2>&1
cat $OUT
