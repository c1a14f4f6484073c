#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "layernorm or LayerNorm or shadow or encoder or transformer or train or oracle" > gpurun_out/ln_tests.log 2>&1
echo "ln tests rc=$?"; tail -4 gpurun_out/ln_tests.log
for t in 1 0; do
WEEDCU_LN_CLUSTER=$t timeout 600 python tools/microbench.py --group ew --out gpurun_out/mb_ln$t.json 2>&1 | grep -i "layernorm" 
done
