#!/bin/bash
# validate: head-grouping fix, decode kernels, skinny matmul, Adam-refreshed weight shadows; decode bench; launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
tail -n 1 gpurun_out/bench_full.log | cut -c1-300
timeout 600 python tools/decode_bench.py > gpurun_out/decode_fp32.log 2>&1
tail -n 2 gpurun_out/decode_fp32.log | cut -c1-600
timeout 600 python tools/decode_bench.py --precision bf16 > gpurun_out/decode_bf16.log 2>&1
tail -n 1 gpurun_out/decode_bf16.log | cut -c1-400
timeout 600 python tools/decode_bench.py --fused 0 --new 32 > gpurun_out/decode_unfused.log 2>&1
tail -n 1 gpurun_out/decode_unfused.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_launches_v3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/bench_ncu.log 2>&1
tail -1 gpurun_out/bench_ncu.log | cut -c1-100
