#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests_08.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r02_gpu_tests_08.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_08.json 2> gpurun_out/r02_bench_08.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_08.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_08.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config']['loss_first'], d['config']['loss_last'], d['gpu_launches'])
print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernel_breakdown'].items()})
PY
