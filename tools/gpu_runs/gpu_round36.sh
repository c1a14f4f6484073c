#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 700 --csv --log-file gpurun_out/r01_decode_launches_v12.csv python tools/decode_bench.py --new 32 > gpurun_out/decode_ncu.log 2>&1
echo rc=$?; tail -1 gpurun_out/decode_ncu.log | cut -c1-200
