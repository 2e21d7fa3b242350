#!/bin/bash
# ncu: launch list of a bench run + full captures of the stage kernels (traffic, pipe utilisation)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stage_fused|node_update|estab|deltat|spmv2" -s 60 -c 8 -o gpurun_out/r2_stage -f python tools/exp_stage.py 2829 2>&1 | tail -2
ls -la gpurun_out/r2_stage.ncu-rep gpurun_out/r2_launches.csv
