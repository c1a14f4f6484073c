#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_host_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/t_host.log
timeout 600 python tools/microbench.py --group tc --out gpurun_out/mb_tc.json > gpurun_out/mb_tc.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --layers 2 --no-cpu-baseline > gpurun_out/bench_l2.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --layers 2 --no-cpu-baseline --no-clocks > gpurun_out/bench_l2_noclk.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
for f in t_kernels t_host mb_tc bench_l2 bench_l2_noclk bench_full; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-400; done
