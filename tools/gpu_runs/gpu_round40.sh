#!/bin/bash
# 2 GPUs: sensitivity of the overlapped gradient all-reduce to the bucket size
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for bb in 8388608 33554432 134217728; do
  WH_DP_BUCKET_BYTES=$bb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_bb$bb.json 2> gpurun_out/n2_bb$bb.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n2_bb$bb.json').read().strip().splitlines()[-1]); print('bucket $bb', round(d['ms_per_step'],3), round(d['value'],1))
except Exception as e: print('bucket $bb no result', e)
PY
done
WH_DP_OVERLAP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_after.json 2> gpurun_out/n2_after.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n2_after.json').read().strip().splitlines()[-1]); print('after-backward', round(d['ms_per_step'],3), round(d['value'],1))
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n1_same_box.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/n1_same_box.json').read().strip().splitlines()[-1]); print('n1', round(d['ms_per_step'],3), round(d['value'],1))
PY
