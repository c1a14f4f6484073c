#!/bin/bash
# last check of the session: full GPU suite + smoke + a short bench on the final code
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_last.log 2>&1
echo "gpu tests rc=$?"; tail -2 gpurun_out/gpu_tests_last.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_last.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r01_bench_last.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config']['loss_last'], d['gpu_launches'])
PY
