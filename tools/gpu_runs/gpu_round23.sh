#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "8192 50257 768 1 0 0 1256001 head_pair_v11" "8192 50257 768 1 0 0 256201 head_single_v11"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 --launch-skip 3 -c 1 -f -o gpurun_out/r01_gemm_$8 python tools/gemm_probe.py $1 $2 $3 $4 $5 $6 $7 > gpurun_out/probe_$8.log 2>&1
  tail -1 gpurun_out/probe_$8.log
done
