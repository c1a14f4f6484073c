#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm or bf16 or matmul" > gpurun_out/r02_gemm_tests_06.log 2>&1
echo "gemm tests rc=$?"; tail -8 gpurun_out/r02_gemm_tests_06.log
timeout 600 python tools/epilogue_bench.py > gpurun_out/r02_epilogue_bench_06.log 2>&1
echo rc=$?; cat gpurun_out/r02_epilogue_bench_06.log | tail -30
