#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs > gpurun_out/r02_bench_22.json 2> gpurun_out/r02_bench_22.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_22.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config'].get('loss_last'), d['gpu_launches'], d['roofline'])
PY
