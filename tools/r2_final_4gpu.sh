#!/bin/bash
# 4-GPU box: the bench at N=4 (weak scaling, parity self-check on 4 ranks, 64 M strong-scaling domain), final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
tail -c 300 gpurun_out/r2_bench_n4.json; tail -2 gpurun_out/r2_bench_n4.err
