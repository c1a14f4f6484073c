#!/bin/bash
# launch list (ncu gpu__time_duration) of ~2 steps, current code
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1900 -c 700 --csv --log-file gpurun_out/r02_launches_24.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-clocks --no-configs > gpurun_out/r02_launches_24.log 2>&1
echo rc=$?; tail -2 gpurun_out/r02_launches_24.log; wc -l gpurun_out/r02_launches_24.csv
