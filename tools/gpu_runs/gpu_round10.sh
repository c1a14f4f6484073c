#!/bin/bash
# validate: streaming LayerNorm passes, vectorised heads relayout, branch-free flash softmax, short-axis reduce
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
tail -n 1 gpurun_out/bench_full.log | cut -c1-300
timeout 600 python tools/microbench.py --group attn --out gpurun_out/microbench_attn.json > gpurun_out/microbench_attn.log 2>&1
cat gpurun_out/microbench_attn.log
timeout 900 python tools/microbench.py --group ew --out gpurun_out/microbench_ew.json > gpurun_out/microbench_ew.log 2>&1
grep -E "layernorm|reduce" gpurun_out/microbench_ew.log
