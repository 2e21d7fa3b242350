#!/bin/bash
cd "$GRAFT_REPO_ROOT"
OUT=gpurun_out/exp_r1h.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/exp_stage.py 2829 >> $OUT 2>&1; }
AB=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_ab.so
run CFDB_TILE=1 CFDB_TILE_MINB=4
run CFDB_TILE=1 CFDB_TILE_MINB=4 CFDB_LIB_PATH=$AB
run CFDB_TILE=1 CFDB_LIB_PATH=$AB
CFDB_TILE=1 CFDB_TILE_MINB=4 CFDB_LIB_PATH=$AB python tests/opt_worker.py >> $OUT 2>&1
cat $OUT | cut -c1-170
