#!/bin/bash
# the default bench command on the final code
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 100 python bench.py > gpurun_out/r01_bench_full_v18.json 2> gpurun_out/bench_v18.err
echo "bench rc=$?"; cut -c1-220 gpurun_out/r01_bench_full_v18.json
