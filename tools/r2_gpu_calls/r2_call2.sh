#!/bin/bash
# round 2, GPU call 2: full GPU suite on a 2-GPU box (multi-GPU parity included), bench at N=1 and N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_gputests.log
cat gpurun_out/r2_gputests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
tail -c 3000 gpurun_out/r2_bench_n1_a.json; tail -5 gpurun_out/r2_bench_n1_a.err
CFDB_NO_GRAPH=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2_bench_n1_nograph.json 2> gpurun_out/r2_bench_n1_nograph.err
tail -c 600 gpurun_out/r2_bench_n1_nograph.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err
tail -c 1500 gpurun_out/r2_bench_n2_a.json; tail -5 gpurun_out/r2_bench_n2_a.err
