#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for m in 0 1 2; do
WEEDCU_GEMM_MODE=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks > gpurun_out/r02_mode$m.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_mode$m.json').read().strip().splitlines()[-1])
print('mode $m', round(d['ms_per_step'],3), round(d['kernel_breakdown']['gemm_bf16_tcgen05']['ms_per_step'],3))
PY
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
