#!/bin/bash
# 8-GPU box: the bench at N=8 (weak scaling, parity self-check on 8 ranks, 64 M strong-scaling domain)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 1500 gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err
