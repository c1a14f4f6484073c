#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:flash_attn_fwd -s 4 -c 1 -f -o gpurun_out/r02_flash python tools/microbench.py --group attn > gpurun_out/r02_flash_ncu.log 2>&1
echo "ncu flash rc=$?"; tail -3 gpurun_out/r02_flash_ncu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ce_fwd_partial_bf16 -s 2 -c 1 -f -o gpurun_out/r02_cefwd python tools/ce_bench.py > gpurun_out/r02_cefwd_ncu.log 2>&1
echo "ncu ce rc=$?"; tail -3 gpurun_out/r02_cefwd_ncu.log
timeout 300 python tools/microbench.py --group attn 2>&1 | tail -5
