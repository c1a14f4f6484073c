#!/bin/bash
# skinny matmul v2 + single-launch LayerNorm: kernel tests, decode bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 600 python tools/decode_bench.py > gpurun_out/decode_fp32.log 2>&1
tail -n 1 gpurun_out/decode_fp32.log | cut -c1-1500
timeout 600 python tools/decode_bench.py --precision bf16 > gpurun_out/decode_bf16.log 2>&1
tail -n 1 gpurun_out/decode_bf16.log | cut -c1-400
