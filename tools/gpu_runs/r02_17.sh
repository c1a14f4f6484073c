#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for st in 1 3; do
CHECK_DP_VERBOSE=1 CHECK_DP_STEPS=$st timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$st bench.py --gpus 2 --check-dp --batch 4 > gpurun_out/r02_check_dp_s$st.json 2> gpurun_out/r02_check_dp_s$st.err
echo "steps=$st rc=$?"
grep "^fp32\|^bf16" gpurun_out/r02_check_dp_s$st.err | cut -c1-1800
done
