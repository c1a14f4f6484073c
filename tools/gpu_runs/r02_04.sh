#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/epilogue_bench.py > gpurun_out/r02_epilogue_bench_04.log 2>&1
echo rc=$?; cat gpurun_out/r02_epilogue_bench_04.log | tail -30
