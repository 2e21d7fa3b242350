#!/bin/bash
# 2 GPUs: ghost refresh only after the last stage where stages 1-3 need none: multi-GPU parity tests, bench N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['multi_gpu_parity'], d['gpu_launches'])
P
