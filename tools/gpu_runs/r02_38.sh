#!/bin/bash
# device timeline of the step's phases at N = 1 and N = 2 (WH_DP_EVENTS=1)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
WH_DP_EVENTS=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks 2>&1 >/dev/null | grep "wh events" | tail -4
for cfg in "A=1" "WH_DP_OVERLAP=0" "WH_DP_SPLIT_ADAM=1"; do
echo "N=2 $cfg"
env $cfg WH_DP_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks 2>&1 >/dev/null | grep "wh events" | tail -3
done
