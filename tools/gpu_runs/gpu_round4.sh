#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/t_host.log
timeout 900 python bench.py --steps 3 --warmup 3 --layers 2 --no-cpu-baseline > gpurun_out/bench_l2.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --layers 2 --no-cpu-baseline --no-clocks > gpurun_out/bench_l2_noclk.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-clocks > gpurun_out/bench_full_noclk.log 2>&1
for f in t_host bench_l2 bench_l2_noclk bench_full bench_full_noclk; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-300; done
