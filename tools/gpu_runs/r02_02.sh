#!/bin/bash
# round 2, call 2: extended GEMM epilogue kernel tests + the host tests fixed since call 1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm_bf16_ex or grouped_bf16out" > gpurun_out/r02_ex_tests.log 2>&1
echo "ex tests rc=$?"; tail -25 gpurun_out/r02_ex_tests.log
timeout 900 python -m pytest tests/test_host_gpu.py -q -m gpu -k "arbiter or c4_scaled or pdl or view_of or c2_tabular_mlp_full" > gpurun_out/r02_host_tests_02.log 2>&1
echo "host tests rc=$?"; tail -8 gpurun_out/r02_host_tests_02.log
