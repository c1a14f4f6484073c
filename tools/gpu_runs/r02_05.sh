#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 1 -c 4 -o gpurun_out/r02_epilogue_prof -f python tools/epilogue_prof.py > gpurun_out/r02_epilogue_prof.log 2>&1
echo rc=$?; tail -5 gpurun_out/r02_epilogue_prof.log; ls -la gpurun_out/*.ncu-rep
