#!/bin/bash
# dynamic tile scheduler in the CTA-pair GEMM: correctness (kernel + host suites with it on) and single-GPU speed
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
WEEDCU_GEMM_DYNAMIC=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or bf16 or pair or epilogue or matmul" 2>&1 | tail -3
WEEDCU_GEMM_DYNAMIC=1 timeout 600 python -m pytest tests/test_host_gpu.py -m gpu -x -q 2>&1 | tail -3
for d in 0 1 0 1; do
WEEDCU_GEMM_DYNAMIC=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks > gpurun_out/r02_dyn$d.json 2>gpurun_out/r02_dyn$d.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_dyn$d.json').read().strip().splitlines()[-1])
    print('dynamic=$d', round(d['ms_per_step'],3), round(d['value'],1), d['config']['loss_last'])
except Exception as e: print('ERR', e); print(open('gpurun_out/r02_dyn$d.err').read()[-600:])
PY
done
