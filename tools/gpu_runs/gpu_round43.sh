#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "layernorm or LayerNorm or encoder or cow or train" > gpurun_out/ln_tests.log 2>&1
echo "ln tests rc=$?"; tail -3 gpurun_out/ln_tests.log
timeout 300 python tools/microbench.py --group ew --out gpurun_out/mb.json 2>&1 | grep -i "layernorm"
