#!/bin/bash
# Host-library parity tests + smoke + first full bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/t_kernels.log
timeout 1500 python -m pytest tests/test_host_gpu.py -m gpu -q -x --maxfail=200 2>&1 | tail -120 > gpurun_out/t_host.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --layers 2 --no-cpu-baseline > gpurun_out/bench_l2.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
for f in t_kernels t_host smoke bench_l2 bench_full; do echo "== $f"; tail -n 6 gpurun_out/$f.log | cut -c1-600; done
