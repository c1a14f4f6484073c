#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_gpu.py -m gpu -x -q -s -k "bf16_vs_fp32 or cow or deferred or view_of or operand_cache" 2>&1 | grep -v "^$" | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs > gpurun_out/r02_bench_33.json 2> gpurun_out/r02_bench_33.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_33.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])
PY
