#!/bin/bash
# DRAM bytes per GEMM launch of one step (roofline.traffic) for the round-2 code
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 --launch-skip 333 -c 111 --csv --log-file gpurun_out/r02_gemm_dram_v35.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks --no-parity --no-configs > gpurun_out/r02_gemm_dram_v35.log 2>&1
echo rc=$?; wc -l gpurun_out/r02_gemm_dram_v35.csv
