#!/bin/bash
# final single-GPU measurements of round 2: launch list under ncu, full captures of the stage kernels, the bench line, the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-secondary --no-parity > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stage_fused|boundary_update|estab|deltat|spmv2" -s 60 -c 8 -o gpurun_out/r2_stage -f python tools/exp_stage.py 2829 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2_bench_ref.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
