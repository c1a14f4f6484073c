#!/bin/bash
# fused CE backward pack in the step, CE/LN microbench (LN single-launch experiment), decode re-check
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
tail -n 1 gpurun_out/bench_full.log | cut -c1-200
WEEDCU_LN_SINGLE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lnsingle.log 2>&1
tail -n 1 gpurun_out/bench_lnsingle.log | cut -c1-200
timeout 600 python tools/microbench.py --group ew --out gpurun_out/microbench_ew.json > gpurun_out/microbench_ew.log 2>&1
grep -E "layernorm|cross_entropy|softmax" gpurun_out/microbench_ew.log
WEEDCU_LN_SINGLE=1 timeout 600 python tools/microbench.py --group ew --out gpurun_out/microbench_ew_lnsingle.json > gpurun_out/microbench_ew_lnsingle.log 2>&1
grep -E "layernorm_fwd" gpurun_out/microbench_ew_lnsingle.log
timeout 600 python tools/decode_bench.py > gpurun_out/decode_fp32.log 2>&1
tail -n 1 gpurun_out/decode_fp32.log | cut -c1-1300
