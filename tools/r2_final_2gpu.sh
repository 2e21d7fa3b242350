#!/bin/bash
# 2-GPU box, final build: full GPU suite (multi-GPU parity included), bench at N=2 (parity self-check, strong-scaling secondary), reference arm under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_gputests_2gpu.txt; cat gpurun_out/r2_gputests_2gpu.txt | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 1800 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
