#!/bin/bash
# final evidence run: tests, bench (with cpu baseline + reference arm), microbench, decode, ncu launch list + full captures
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1
tail -n 1 gpurun_out/bench_default.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
tail -n 1 gpurun_out/bench_reference.log | cut -c1-300
timeout 900 python tools/microbench.py > gpurun_out/microbench.log 2>&1
tail -3 gpurun_out/microbench.log
timeout 600 python tools/decode_bench.py > gpurun_out/decode_fp32.log 2>&1
tail -n 1 gpurun_out/decode_fp32.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/bench_ncu.log 2>&1
tail -1 gpurun_out/bench_ncu.log | cut -c1-100
