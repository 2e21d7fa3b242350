#!/bin/bash
# end-of-round measurement on one B200: bench line, reference arm, ncu launch list, ncu full capture of the step's kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"calcrhs_elem|node_update|estab|deltat" \
    --launch-skip 20 --launch-count 10 -f -o gpurun_out/prof_stage_r1_final \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_r1.json gpurun_out/bench_r1_ref.json
tail -3 gpurun_out/bench_r1.err
ls -la gpurun_out/*.ncu-rep
