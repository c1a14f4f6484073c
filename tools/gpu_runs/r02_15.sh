#!/bin/bash
# 2 GPUs: --check-dp with per-parameter Adam-moment gate
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --check-dp > gpurun_out/r02_check_dp_n2.json 2> gpurun_out/r02_check_dp_n2.err
echo "check-dp rc=$?"; tail -1 gpurun_out/r02_check_dp_n2.json | cut -c1-3000; tail -3 gpurun_out/r02_check_dp_n2.err
