#!/bin/bash
# staged SpMV kernels + NVML clock sampler: parity suite, per-kernel timings, the bench line with its secondary configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c21_tests.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_c21_tests.txt
timeout 300 python tools/exp_stage.py 2829 2>&1 | tail -1
timeout 900 python bench.py --no-cpu > gpurun_out/r2_c21_bench.json 2> gpurun_out/r2_c21_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_c21_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_c21_bench.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['clocks'])
print(json.dumps(d['config']['secondary'].get('ale4M'))[:900])
P
