#!/bin/bash
# correctness of the reworked kernels + microbench + full bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 900 python tools/microbench.py > gpurun_out/microbench.log 2>&1
grep -E "softmax|layernorm|cross_entropy|adam|pack|unary" gpurun_out/microbench.log | head -40
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
tail -n 1 gpurun_out/bench_full.log | cut -c1-400
