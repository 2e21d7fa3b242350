#!/bin/bash
# GPU batch: new reference-pinned parity tests, then stage-overlap / co-residency experiments on the bench mesh
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_pin.py -x -q -m gpu > gpurun_out/refpin_gpu.log 2>&1
echo "refpin exit $?" >> gpurun_out/refpin_gpu.log
OUT=gpurun_out/exp_r1c.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/exp_stage.py 2829 >> $OUT 2>&1; }
run A=baseline
run CFDB_STAGE_OVERLAP=1
run CFDB_CALCRHS_PAD_KB=60
run CFDB_STAGE_OVERLAP=1 CFDB_CALCRHS_PAD_KB=60
run CFDB_STAGE_OVERLAP=1 CFDB_CALCRHS_PAD_KB=100
run CFDB_STAGE_OVERLAP=1 CFDB_CALCRHS_PAD_KB=50 CFDB_CALCRHS_MINB=5
run CFDB_STAGE_OVERLAP=1 CFDB_CALCRHS_PAD_KB=60 CFDB_CALCRHS_MINB=5
cat $OUT
tail -3 gpurun_out/refpin_gpu.log
