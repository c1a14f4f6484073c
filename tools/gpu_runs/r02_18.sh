#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "cross_entropy or epilogue" 2>&1 | tail -5
timeout 300 python tools/ce_bench.py 2>&1 | tail -4
