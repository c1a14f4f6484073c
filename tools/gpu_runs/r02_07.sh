#!/bin/bash
# launch list (ncu gpu__time_duration) of 1 step with the epilogue fusions on
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1900 -c 520 --csv --log-file gpurun_out/r02_launches_07.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-clocks > gpurun_out/r02_launches_07.log 2>&1
echo rc=$?; tail -2 gpurun_out/r02_launches_07.log; wc -l gpurun_out/r02_launches_07.csv
