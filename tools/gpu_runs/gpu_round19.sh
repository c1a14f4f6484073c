#!/bin/bash
# CTA-pair (cta_group::2) GEMM: parity tests first, then the tile-configuration sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "cta_pair or grouped_equals" > gpurun_out/pair_tests.log 2>&1
echo "pair tests rc=$?"; tail -15 gpurun_out/pair_tests.log
timeout 600 python tools/microbench.py --group tune --out gpurun_out/r01_tune_v9.json > gpurun_out/r01_tune_v9.log 2>&1
echo "tune rc=$?"; grep -E "BEST|mode        [012] " gpurun_out/r01_tune_v9.log | tail -60
