#!/bin/bash
# re-entry state check: full GPU tests, bench (flash on/off), microbench, ncu launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>&1
tail -n 1 gpurun_out/bench_full.log | cut -c1-400
WEEDCU_FLASH=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noflash.log 2>&1
tail -n 1 gpurun_out/bench_noflash.log | cut -c1-300
timeout 900 python tools/microbench.py > gpurun_out/microbench.log 2>&1
tail -5 gpurun_out/microbench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -2 gpurun_out/bench_ncu.log | cut -c1-200
