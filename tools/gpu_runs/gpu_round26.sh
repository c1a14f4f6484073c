#!/bin/bash
# ncu launch list of the bench command (per-launch gpu time, cold-cache and serialised)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1050 -c 700 --csv --log-file gpurun_out/r01_launches_v15.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/bench_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/bench_ncu.log | cut -c1-300
