#!/bin/bash
# A/B: default library against cfd_b200/libcfdb200_ab*.so builds (csrc/Makefile: ab), per-kernel timings on the bench mesh
cd "$GRAFT_REPO_ROOT"
OUT=gpurun_out/exp_ab.txt
: > $OUT
echo "== default" >> $OUT; timeout 300 python tools/exp_stage.py 2829 >> $OUT 2>&1
for lib in cfd_b200/libcfdb200_ab*.so; do
  echo "== $lib" >> $OUT; CFDB_LIB_PATH=$GRAFT_REPO_ROOT/$lib timeout 300 python tools/exp_stage.py 2829 >> $OUT 2>&1
done
echo "== default" >> $OUT; timeout 300 python tools/exp_stage.py 2829 >> $OUT 2>&1
cut -c1-140 $OUT
