#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for d in 0 1 0 1; do
WEEDCU_GEMM_DYNAMIC=$d WH_DP_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline --no-parity --no-configs --no-clocks 2> gpurun_out/r02_43.err > gpurun_out/r02_43.json
grep "wh events" gpurun_out/r02_43.err | tail -1
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_43.json').read().strip().splitlines()[-1]); print('N=2 dynamic=$d', round(d['ms_per_step'],3), round(d['value'],1))
PY
done
