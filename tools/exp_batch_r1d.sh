#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "topology or optional_paths or reference_program or golden" > gpurun_out/topo_gpu_tests.log 2>&1
echo "exit $?" >> gpurun_out/topo_gpu_tests.log
OUT=gpurun_out/exp_r1d.txt
: > $OUT
python tools/exp_setup.py 2829 >> $OUT 2>&1
CFDB_HOST_TOPO=1 python tools/exp_setup.py 2829 >> $OUT 2>&1
python tools/exp_setup.py 5657 >> $OUT 2>&1
CFDB_HOST_TOPO=1 python tools/exp_setup.py 5657 >> $OUT 2>&1
cat $OUT
tail -5 gpurun_out/topo_gpu_tests.log
