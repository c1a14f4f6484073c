#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
WH_DP_SPLIT_ADAM=1 WH_DP_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks 2>&1 >/dev/null | grep "wh split\|wh events" | tail -6
