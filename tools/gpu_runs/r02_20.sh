#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "attention or attn or flash or mha or encoder or gpt" 2>&1 | tail -5
timeout 300 python tools/microbench.py --group attn 2>&1 | tail -5
