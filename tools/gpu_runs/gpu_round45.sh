#!/bin/bash
# 8 GPUs, driver-style launch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r01_bench_n8_v17.json 2> gpurun_out/n8.err
echo "n8 rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r01_bench_n8_v17.json').read().strip().splitlines()[-1]); print('n8', round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['config']['loss_last'])
PY
tail -2 gpurun_out/n8.err
