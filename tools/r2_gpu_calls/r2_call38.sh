#!/bin/bash
# boundary_update with eight records in flight: register budgets
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== minb6 (80 regs)"; timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-130
echo "== minb8 (64 regs)"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_bnd8.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-130
echo "== minb10 (48 regs)"; CFDB_LIB_PATH=cfd_b200/libcfdb200_ab_bnd10.so timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1 | cut -c1-130
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c38_tests.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_c38_tests.txt
