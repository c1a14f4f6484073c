#!/bin/bash
# 4 staging buffers in the GEMM epilogue: parity, then the tile-configuration sweep again
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm or bf16 or matmul" > gpurun_out/gemm_tests.log 2>&1
echo "gemm tests rc=$?"; tail -5 gpurun_out/gemm_tests.log
timeout 600 python tools/microbench.py --group tune --out gpurun_out/r01_tune_v10.json > gpurun_out/r01_tune_v10.log 2>&1
echo "tune rc=$?"; grep -E "BEST|mode        [012] " gpurun_out/r01_tune_v10.log | tail -60
