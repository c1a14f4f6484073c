#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "layernorm or epilogue or gpt" 2>&1 | tail -4
timeout 600 python tools/microbench.py --group ew 2>&1 | grep -i "layernorm\|gelu\|adam\|pack\|softmax" | head -40
