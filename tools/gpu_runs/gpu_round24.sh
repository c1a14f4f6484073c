#!/bin/bash
# full GPU suite + C5 bench with the CTA-pair GEMM family under the calibrated cost model
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_full_v9.json 2> gpurun_out/bench_v9.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01_bench_full_v9.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'])
for k,v in d['kernel_breakdown'].items(): print(k, round(v['ms_per_step'],3), v['launches_per_step'])
print(d['clocks'], d['host'])
PY
