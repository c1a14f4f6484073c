#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 300 python tools/microbench.py --group ew --out gpurun_out/mb.json 2>&1 | grep -i "cross_entropy"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r01_bench_full_v17.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r01_bench_full_v17.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config']['loss_last'])
PY
