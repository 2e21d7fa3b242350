#!/bin/bash
# double-buffered device staging of cfdb_step_streamed: parity of the streamed path, e2e figure
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "streamed" 2>&1 | tail -2
timeout 900 python bench.py --no-cpu --no-secondary --no-parity > gpurun_out/r2_c40_bench.json 2> gpurun_out/r2_c40_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_c40_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['steps'])
P
