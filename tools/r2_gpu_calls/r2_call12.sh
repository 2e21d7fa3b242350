#!/bin/bash
# 2-GPU box: full GPU suite (multi-GPU parity included), bench at N=2 with the parity self-check and the strong-scaling secondary
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_gputests12.log; cat gpurun_out/r2_gputests12.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err
tail -c 2500 gpurun_out/r2_bench_n2_b.json; tail -5 gpurun_out/r2_bench_n2_b.err
