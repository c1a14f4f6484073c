#!/bin/bash
# First GPU contact: kernel parity tests, then microbenchmarks. Each step under its own timeout
# and in its own process so a trapped kernel cannot take the rest down.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "not bf16" 2>&1 | tail -60 > gpurun_out/t_kernels.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "bf16" 2>&1 | tail -80 > gpurun_out/t_bf16.log
timeout 900 python tools/microbench.py --group ew --out gpurun_out/mb_ew.json > gpurun_out/mb_ew.log 2>&1
timeout 300 python tools/microbench.py --group f32 --out gpurun_out/mb_f32.json > gpurun_out/mb_f32.log 2>&1
timeout 600 python tools/microbench.py --group tc --out gpurun_out/mb_tc.json > gpurun_out/mb_tc.log 2>&1
for f in gpurun_out/t_kernels.log gpurun_out/t_bf16.log; do tail -n 4 $f; done
