#!/bin/bash
# 8 GPUs: phase timeline with and without the Adam split around the last flushed bucket
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "WH_DP_SPLIT_ADAM=1" "WH_DP_SPLIT_ADAM=0"; do
echo "N=8 $cfg"
env $cfg WH_DP_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 8 --steps 12 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks 2> gpurun_out/r02_41.err > gpurun_out/r02_41_$cfg.json
grep "wh split\|wh events" gpurun_out/r02_41.err | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_41_$cfg.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value'],1))
PY
done
