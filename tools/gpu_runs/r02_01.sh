#!/bin/bash
# round 2, call 1: full GPU suite (new B>1 parity tests, PDL on/off, C2 full rows, C4 scaled), smoke, default bench with parity legs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests_01.log 2>&1
echo "gpu tests rc=$?"; tail -15 gpurun_out/r02_gpu_tests_01.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_01.json 2> gpurun_out/r02_bench_01.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_01.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_01.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config']['loss_last'], d['gpu_launches']); print(d['config']['parity']); print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernel_breakdown'].items()})
PY
