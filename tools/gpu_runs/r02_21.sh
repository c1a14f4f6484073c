#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:flash_attn_fwd -s 4 -c 1 -f -o gpurun_out/r02_flash2 python tools/microbench.py --group attn > gpurun_out/r02_flash2_ncu.log 2>&1
echo "ncu flash rc=$?"; tail -2 gpurun_out/r02_flash2_ncu.log
