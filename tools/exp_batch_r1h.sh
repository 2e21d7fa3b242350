#!/bin/bash
# pathological-operand parity, rolled Gauss-point loop A/B, one full ncu capture of calcrhs_elem and estab
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1h_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/r1h_pytest.log | tail -3
OUT=gpurun_out/exp_r1h.txt
: > $OUT
run() { echo "== $1" >> $OUT; shift; env "$@" timeout 300 python tools/exp_stage.py 2829 $VISC >> $OUT 2>&1; }
VISC=
run default X=1
run rollk CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_ab_rollk.so
run rollk-nb1 CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_ab_rollk.so CFDB_CALCRHS_NB=1
run rollk-minb3 CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_ab_rollk.so CFDB_CALCRHS_MINB=3
run rollk-minb5 CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_ab_rollk.so CFDB_CALCRHS_MINB=5
run default X=1
VISC=visc
run visc-default X=1
run visc-rollk CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_ab_rollk.so
run visc-rollk-minb4 CFDB_LIB_PATH=$GRAFT_REPO_ROOT/cfd_b200/libcfdb200_ab_rollk.so CFDB_CALCRHS_MINB=4
cut -c1-150 $OUT
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"calcrhs_elem|estab" \
    --launch-skip 12 --launch-count 3 -f -o gpurun_out/prof_r1h \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_r1h.log 2>&1
ls -la gpurun_out/*.ncu-rep
