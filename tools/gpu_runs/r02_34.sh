#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "softmax" 2>&1 | tail -3
timeout 600 python tools/microbench.py --group ew 2>&1 | grep -i "softmax" | head -20
