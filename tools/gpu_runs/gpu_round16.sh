#!/bin/bash
# producer-side bf16 shadows (LayerNorm apply, GELU): tests + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
tail -n 1 gpurun_out/bench_full.log | cut -c1-200
