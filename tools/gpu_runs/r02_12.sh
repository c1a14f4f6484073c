#!/bin/bash
# split-K fp32 GEMM: full GPU suite + per-class times of C2/C4 + check-dp style single-GPU sweep of c2/c4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/config_prof.py 2>&1 | tail -12
